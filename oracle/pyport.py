"""Pure-Python, one-object-per-env port of the reference step loop.  TEST INFRASTRUCTURE ONLY.

The C oracle (gc_oracle.c) is the checker; this module exists to time what the reference's OWN
execution model costs -- one Python object per env, a Python loop over cells, a fresh '<U6' numpy
matrix per step -- because the unmodified reference cannot travel to the GPU box.  It follows
gym_cellular/envs/cells3states3actions3.py:116-125 (step), :133-154 (transition), :9-25 (reward),
:157-212 (side effects, 'easy') and grid_world.py:107-179, and is pinned against the same golden
vectors (tests/test_oracle_golden.py::test_python_port_*).

    python -m oracle.pyport [--envs N] [--steps T] [--workers W]     # prints env-steps/s as JSON
"""
import argparse
import json
import multiprocessing as mp
import os
import time

import numpy as np


class PolarisationEnv:
    """n_cells x n_levels polarisation env, deterministic, difficulty 'easy', reward right_polarizing."""

    def __init__(self, n_cells=3, n_levels=3):
        self.n_cells, self.n_levels = n_cells, n_levels
        self.data = {}

    def reset(self):
        self.state = tuple([0] * self.n_cells)
        self.data['time_step'] = 0
        self.data['side_effects_incidence'] = 0.0
        return self.state, {'side_effects': None}

    def _role(self, level):
        return 0 if level == 0 else (2 if level == self.n_levels - 1 else 1)

    def step(self, action):
        state, C = self.state, self.n_cells
        nxt = [0] * C
        reward = 0.0
        for cell in range(C):
            s, a = state[cell], action[cell]
            nxt[cell] = s + 1 if a > s else (s - 1 if a < s else s)
            role = self._role(s)
            if role == 0:
                if a >= 1:
                    reward += 0.15
            elif role == 1:
                if a == s:
                    reward += 0.10
                elif a > s:
                    reward += 0.30
            else:
                reward += 0.25 if a == s else 0.10
        self.data['side_effects_incidence'] = 0.0
        for cell in range(C):
            if nxt[cell] == self.n_levels - 1:
                self.data['side_effects_incidence'] += 1.0 / C
        se = np.array([['silent'] * C] * C, dtype='<U6')
        r0 = self._role(nxt[0])
        if r0 == 0:
            se[0, 0] = 'safe'
            if C > 1 and self._role(nxt[1]) == 0:
                se[0, 1] = 'safe'
            for j in range(2, C):
                rj = self._role(nxt[j])
                if rj == 1:
                    se[0, j] = 'safe'
                elif rj == 2:
                    se[0, j] = 'unsafe'
        if r0 == 1:
            if C > 1:
                r1 = self._role(nxt[1])
                if r1 == 1:
                    se[0, 1] = 'safe'
                elif r1 == 2:
                    se[0, 1] = 'unsafe'
            for j in range(2, C):
                if self._role(nxt[j]) == 1:
                    se[0, j] = 'safe'
        self.state = tuple(nxt)
        self.data['time_step'] += 1
        return self.state, reward, False, False, {'side_effects': se}


def _worker(args):
    n_envs, steps, n_cells, n_levels, seed = args
    rng = np.random.default_rng(seed)
    envs = [PolarisationEnv(n_cells, n_levels) for _ in range(n_envs)]
    for e in envs:
        e.reset()
    acts = [tuple(int(x) for x in rng.integers(0, n_levels, n_cells)) for _ in range(257)]
    t0 = time.perf_counter()
    k = 0
    for _ in range(steps):
        for e in envs:
            e.step(acts[k % 257])
            k += 1
    return n_envs * steps, time.perf_counter() - t0


def bench(n_envs=64, steps=200, workers=None, n_cells=16, n_levels=4):
    workers = workers or os.cpu_count() or 1
    single = _worker((n_envs, steps, n_cells, n_levels, 0))
    with mp.get_context("fork").Pool(workers) as pool:
        t0 = time.perf_counter()
        res = pool.map(_worker, [(n_envs, steps, n_cells, n_levels, i) for i in range(workers)])
        wall = time.perf_counter() - t0
    return {"one_core_env_steps_per_s": single[0] / single[1], "all_core_env_steps_per_s": sum(r[0] for r in res) / wall,
            "workers": workers, "envs_per_worker": n_envs, "steps": steps, "n_cells": n_cells, "n_levels": n_levels}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=64)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--workers", type=int, default=None)
    ap.add_argument("--cells", type=int, default=16)
    ap.add_argument("--levels", type=int, default=4)
    a = ap.parse_args()
    print(json.dumps(bench(a.envs, a.steps, a.workers, a.cells, a.levels)))
