#!/usr/bin/env python
"""Host-side cost of one step launch (us per call), measured with a tiny batch; GPU box only."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_cellular_b200 as B   # noqa: E402

env = B.CellularVectorEnv(num_envs=1024, emit_side_effects=False)
a = torch.zeros(3, env.ld, dtype=torch.int8, device="cuda")
call = env.bind_step(a)
dev = env.device


def per_call(f, n=20000):
    for _ in range(200):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / n * 1e6


print("bind_step launch           %.2f us" % per_call(call))
print("step_device(a)             %.2f us" % per_call(lambda: env.step_device(a)))
print("step(a) full API           %.2f us" % per_call(lambda: env.step(a)))
print("current_stream().cuda_stream %.2f us" % per_call(lambda: torch.cuda.current_stream(dev).cuda_stream))
fn, h = env._lib.gc_step_bound, env._h
s = torch.cuda.current_stream(dev).cuda_stream
print("raw gc_step_bound ctypes   %.2f us" % per_call(lambda: fn(h, 0, s)))
print("torch tiny kernel (a.add_) %.2f us" % per_call(lambda: a.add_(0)))
