#!/usr/bin/env python
"""Grid-world step time against batch size, as a CUDA graph of 8 steps (us per step); GPU box only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_cellular_b200 as B   # noqa: E402

for lg in range(16, 25):
    n = 1 << lg
    env = B.CellularVectorEnv(kind="gridworld", num_envs=n, max_episode_steps=128, emit_side_effects=False)
    ring = []
    for _ in range(8):
        a = torch.full((2, env.ld), 4, dtype=torch.int8, device="cuda")
        a[0] = torch.randint(0, 4, (env.ld,), device="cuda").to(torch.int8)
        ring.append(a)
    g = env.capture_steps(ring)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    reps = max(4, min(200, (1 << 27) // n))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    e1.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * 8)
    print(f"2^{lg} envs  {us:8.2f} us/step  {n / us / 1e3:7.1f} G env-steps/s  {26 * n / us / 1e3:7.0f} GB/s", flush=True)
    env.close()
    del env, ring, g
    torch.cuda.empty_cache()
