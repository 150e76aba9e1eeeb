#!/usr/bin/env python
"""A tabular agent on the packed layout: batched Q-learning over thousands of stochastic polarisation envs.

    python examples/packed_tabular_agent.py [--envs 16384] [--steps 400]

A PilotExperimentation-style tabular agent indexes its model by `prior_knowledge.tabularize(state)` and
`tabularize(action)` (cells3states3actions3.py:281-284) and reads the reward and the side-effect report of
`step()` (:116-125).  On the packed layout the env hands out exactly that: the tabular index, the reward and a
flag byte per env, as device tensors, with one kernel launch per step and nothing else on the hot path.  The joint
action goes in as a packed word (2 bits per cell); for three actions per cell the tabular action index (base 3)
is mapped to its word through a 27-entry table built once.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_cellular_b200 as gcb                                      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--unsafe-penalty", type=float, default=2.0)
    args = ap.parse_args()
    n, C, S, A = args.envs, 3, 3, 3
    env = gcb.PackedCellularVectorEnv(num_envs=n, n_cells=C, n_states=S, stochastic=True, env_seed=1, difficulty="easy")
    n_s, n_a = S ** C, A ** C
    # tabular action index -> packed action word (cell 0 least significant in both)
    digits = torch.arange(n_a, device="cuda")
    word_of_action = torch.zeros(n_a, dtype=torch.int32, device="cuda")
    for c in range(C):
        word_of_action |= ((digits // A ** c) % A).to(torch.int32) << (2 * c)
    Q = torch.zeros(n_s, n_a, device="cuda")
    visits = torch.zeros(n_s, n_a, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(0)
    state = env.tabular_state()
    words = torch.zeros(env.ld, dtype=torch.int32, device="cuda")
    total_reward, unsafe_steps, late_unsafe = 0.0, 0, 0
    for t in range(args.steps):
        eps = max(0.05, 1.0 - t / (0.6 * args.steps))
        greedy = Q[state].argmax(dim=1)
        explore = torch.rand(n, device="cuda", generator=gen) < eps
        action = torch.where(explore, torch.randint(0, n_a, (n,), device="cuda", generator=gen), greedy)
        words[:n] = word_of_action[action]
        _, reward, _, _, infos = env.step(words)                      # one kernel launch
        nxt = infos["tabular_state"].to(torch.int64)
        unsafe = infos["unsafe"]
        target = reward - args.unsafe_penalty * unsafe + 0.9 * Q[nxt].max(dim=1).values
        flat = state * n_a + action
        visits.view(-1).index_add_(0, flat, torch.ones(n, device="cuda"))
        Q.view(-1).index_add_(0, flat, (target - Q.view(-1)[flat]) / visits.view(-1)[flat].clamp(min=1.0))
        total_reward += float(reward.sum())
        unsafe_steps += int(unsafe.sum())
        if t >= args.steps * 3 // 4:
            late_unsafe += int(unsafe.sum())
        state = nxt
    stats = env.stats()
    out = {"env_steps": stats["env_steps"], "mean_reward": total_reward / (n * args.steps),
           "unsafe_rate": unsafe_steps / (n * args.steps), "unsafe_rate_last_quarter": late_unsafe / (n * (args.steps - args.steps * 3 // 4)),
           "kernel_launches": env.launch_count, "stats_reward_sum": stats["reward_sum"], "total_reward": total_reward}
    print(out)
    return out


if __name__ == "__main__":
    main()
