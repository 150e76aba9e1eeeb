#!/usr/bin/env python
"""Device-path step timing for a sweep of env shapes (us per step, algorithmic GB/s); GPU box only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_cellular_b200 as B   # noqa: E402

CASES = [dict(n_cells=3, n_states=3), dict(n_cells=3, n_states=3, stochastic=True), dict(n_cells=2, n_states=3),
         dict(n_cells=8, n_states=4), dict(n_cells=8, n_states=4, stochastic=True), dict(n_cells=16, n_states=4),
         dict(n_cells=16, n_states=4, stochastic=True), dict(n_cells=16, n_states=4, emit_side_effects=True),
         dict(n_cells=10, n_states=8), dict(n_cells=10, n_states=8, stochastic=True), dict(n_cells=6, n_states=5), dict(kind="gridworld"), dict(kind="gridworld", max_episode_steps=128),
         dict(n_cells=10, n_states=8, emit_side_effects=True), dict(n_cells=10, n_states=8, stochastic=True, emit_side_effects=True)]
if "--cells" in sys.argv:      # every cell count of the pair-table kernels, deterministic and stochastic
    CASES = [dict(n_cells=c, n_states=4, stochastic=st) for st in (False, True) for c in range(1, 17)]
if "--only" in sys.argv:       # one case by position (for an ncu capture of its kernel)
    CASES = [CASES[int(sys.argv[sys.argv.index("--only") + 1])]]
PACKED = "--packed" in sys.argv
for kw in CASES:
    kw = dict(kw)
    n = 1 << 24
    kw.setdefault("emit_side_effects", False)
    if PACKED:
        env = B.PackedCellularVectorEnv(num_envs=n, **kw)
        a = torch.randint(-2 ** 31, 2 ** 31 - 1, (env.ld,), dtype=torch.int32, device="cuda") if env.n_actions == 4 else \
            env.pack(torch.randint(0, env.n_actions, (env.n_cells, n), dtype=torch.int8, device="cuda"))
    else:
        env = B.CellularVectorEnv(num_envs=n, **kw)
        a = torch.randint(0, min(env.n_actions, 4), (env.n_cells, n), dtype=torch.int8, device="cuda")
    if env.kind == "gridworld":
        a[1] = 4
    call = env.bind_step(a)
    for _ in range(5):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40):
        call()
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / 40
    b = env.hbm_bytes_per_env_step if PACKED else 3 * env.n_cells + 20 + (env.n_cells if kw["emit_side_effects"] else 0)
    print(f"{str(kw):80s} {ms * 1e3:8.1f} us  {b * n / ms / 1e6:7.0f} GB/s  {n / ms / 1e6:7.1f} G env-steps/s", flush=True)
    env.close()
    del env, a, call
    torch.cuda.empty_cache()
