"""Packed layout on one GPU: device-path step time / roofline fraction (25 B per env-step) and the host
path (4 B in, 9 B out per env-step), next to the int8 layout on the same box.

    python scripts/packed_bench.py [--envs N] [--cells C] [--levels S] [--steps K] [--stochastic] [--no-host]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import gym_cellular_b200 as B


def timed(fn, steps, warmup=10):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        fn(i)
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 24)
    ap.add_argument("--cells", type=int, default=16)
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--stochastic", action="store_true")
    ap.add_argument("--no-host", action="store_true")
    ap.add_argument("--no-int8", action="store_true")
    ap.add_argument("--chunk", type=int, default=1 << 20)
    a = ap.parse_args()
    n, C, S = a.envs, a.cells, a.levels
    peak = 6554.2
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out = {"envs": n, "cells": C, "levels": S, "stochastic": a.stochastic}
    gen = torch.Generator(device="cuda").manual_seed(1)
    kw = dict(num_envs=n, n_cells=C, n_states=S, stochastic=a.stochastic, env_seed=0, host_chunk_envs=a.chunk)
    pk = B.PackedCellularVectorEnv(**kw)
    ring8 = [torch.randint(0, S, (C, pk.ld), dtype=torch.int8, device="cuda", generator=gen) for _ in range(8)]
    ringp = []
    for r in ring8:
        w = torch.zeros(pk.ld, dtype=torch.int32, device="cuda")
        w[:n] = pk.pack(r[:, :n])
        ringp.append(w)
    calls = [pk.bind_step(w) for w in ringp]
    ms = timed(lambda i: calls[i % 8](), a.steps)
    bpe = pk.hbm_bytes_per_env_step
    out["packed"] = {"us_per_step": ms * 1e3, "env_steps_per_s": n / (ms * 1e-3), "bytes_per_env_step": bpe,
                     "achieved_gbs": bpe * n / (ms * 1e-3) / 1e9, "frac": bpe * n / (ms * 1e-3) / 1e9 / peak}
    if not a.no_int8:
        i8 = B.CellularVectorEnv(emit_side_effects=False, **kw)
        calls8 = [i8.bind_step(r) for r in ring8]
        ms8 = timed(lambda i: calls8[i % 8](), a.steps)
        b8 = 3 * C + 20
        out["int8"] = {"us_per_step": ms8 * 1e3, "env_steps_per_s": n / (ms8 * 1e-3), "bytes_per_env_step": b8,
                       "achieved_gbs": b8 * n / (ms8 * 1e-3) / 1e9, "frac": b8 * n / (ms8 * 1e-3) / 1e9 / peak}
        i8.close()
        del i8, calls8
    del ring8
    torch.cuda.empty_cache()
    if not a.no_host:
        hb = pk.host_action_buffer
        hb[:] = ringp[0][:n].cpu().numpy().view(np.uint32)
        for _ in range(2):
            pk.step(hb)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        K = 10
        for _ in range(K):
            obs, rew, term, trunc, info = pk.step(hb)
            sink = float(rew[0]) + int(obs[0])
        el = (time.perf_counter() - t0) / K
        h2d, d2h = pk.host_bytes_per_env_step
        out["host_packed"] = {"ms_per_step": el * 1e3, "env_steps_per_s": n / el, "h2d_bytes": h2d * n, "d2h_bytes": d2h * n,
                              "d2h_gbs": d2h * n / el / 1e9, "chunk": a.chunk}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
