#!/usr/bin/env python
"""Plan on the exact model, evaluate with the fused rollout -- everything on one B200.

    python examples/plan_and_rollout.py [--env gridworld|polarisation] [--envs 1048576] [--steps 128]

1. `exact_model` enumerates P[s, a, s'] and R[s, a] with the step kernel (replay mode),
2. value iteration (torch, on the GPU) turns the model into a tabular policy that avoids unsafe
   successor states where it can,
3. `env.rollout(steps, policy)` runs that policy for `steps` steps in every env inside one kernel and
   is compared with uniformly random actions.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_cellular_b200 as gcb                                      # noqa: E402
from gym_cellular_b200.model import exact_model                     # noqa: E402


def value_iteration(P, R, valid, gamma=0.95, iters=300):
    P, R = torch.from_numpy(P).cuda(), torch.from_numpy(R).cuda()
    V = torch.zeros(P.shape[0], dtype=torch.float64, device="cuda")
    for _ in range(iters):
        Q = R + gamma * (P @ V)
        V = Q.max(dim=1).values
    policy = Q.argmax(dim=1).to(torch.int32)
    return policy, V


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="gridworld", choices=["gridworld", "polarisation"])
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=128)
    args = ap.parse_args()
    kind, kw = ("gridworld", {}) if args.env == "gridworld" else ("cellular", dict(stochastic=True))
    P, R, valid = exact_model(kind, **kw)
    if kind == "gridworld":
        R = R.copy()
        R[:, 24] = -1e9                      # the action that names no position is not an action (KeyError in the reference)
    policy, V = value_iteration(P, R, valid)
    results = {}
    for name, pol in (("planned", policy), ("random", None)):
        env = gcb.CellularVectorEnv(kind=kind, num_envs=args.envs, env_seed=1, emit_side_effects=False, **kw)
        ret, unsafe = env.rollout(args.steps, pol)
        results[name] = (float(ret.mean()), float(unsafe.float().mean()), env.stats()["count_sum"] / (args.envs * args.steps))
        env.close()
    for name, (ret, uns, cnt) in results.items():
        print(f"{name:8s} mean return over {args.steps} steps = {ret:8.3f}   unsafe steps per env = {uns:6.3f}   "
              f"mean count per step = {cnt:5.3f}")
    return results


if __name__ == "__main__":
    main()
