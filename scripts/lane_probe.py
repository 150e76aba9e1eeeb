"""How much of the per-launch fixed cost of the launch-bound configs overlaps when the batch is stepped as K
independent chains (K sub-batches, one stream each, gc_step_many per chain)?  Probe for the `lanes` design.

    python scripts/lane_probe.py [--kind gridworld|cellular] [--envs N] [--steps K]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gym_cellular_b200 as B


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="gridworld")
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--chunk", type=int, default=64)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    out = {}
    for K in (1, 2, 4, 8):
        n = a.envs // K // 16 * 16
        envs, slots, streams = [], [], []
        gen = torch.Generator(device=dev).manual_seed(1)
        for k in range(K):
            if a.kind == "gridworld":
                env = B.CellularVectorEnv(kind="gridworld", num_envs=n, env_seed=0, env_id_offset=k * n, max_episode_steps=128,
                                          emit_side_effects=False)
            else:
                env = B.CellularVectorEnv(num_envs=n, env_seed=0, env_id_offset=k * n, emit_side_effects=False)
            ring = []
            for _ in range(8):
                if a.kind == "gridworld":
                    act = torch.full((2, env.ld), 4, dtype=torch.int8, device=dev)
                    jur = torch.randint(0, 2, (env.ld,), device=dev, generator=gen)
                    pos = torch.randint(0, 4, (env.ld,), device=dev, generator=gen).to(torch.int8)
                    act[0] = torch.where(jur == 0, pos, act[0])
                    act[1] = torch.where(jur == 1, pos, act[1])
                else:
                    act = torch.randint(0, 3, (3, env.ld), dtype=torch.int8, device=dev, generator=gen)
                ring.append(act)
            envs.append(env)
            slots.append([env._bind(r) for r in ring])
            streams.append(torch.cuda.Stream(device=dev))
        main_s = torch.cuda.current_stream(dev)

        def run(steps):
            for s in streams:
                s.wait_stream(main_s)
            for c0 in range(0, steps, a.chunk):
                m = min(a.chunk, steps - c0)
                for env, sl, s in zip(envs, slots, streams):
                    env.step_many(sl, m, stream=s)
            for s in streams:
                main_s.wait_stream(s)
        run(64)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main_s)
        run(a.steps)
        e1.record(main_s)
        e1.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / a.steps
        out[K] = {"us_per_step": round(us, 3), "env_steps_per_s": round(n * K / (us * 1e-6) / 1e9, 2)}
        for env in envs:
            env.close()
    print(json.dumps({"kind": a.kind, "envs": a.envs, "chains": out}))


if __name__ == "__main__":
    main()
