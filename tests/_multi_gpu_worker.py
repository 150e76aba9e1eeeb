"""Worker of tests/test_gpu_multi.py: one rank per GPU (torchrun, NCCL).  Each rank steps its shard of a global
batch (stochastic polarisation, grid world and the packed 16 x 4 layout; Philox keyed by GLOBAL env id) and
rank 0 compares every shard, gathered over NCCL, with the same batch stepped whole on its own GPU."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
import torch.distributed as dist

import gym_cellular_b200 as B
from gym_cellular_b200.distributed import StatsReducer, shard_range


def actions_for(kind, n_cells, n_actions, n_global, t, device):
    gen = torch.Generator(device=device).manual_seed(1000 + t)          # the same global action stream on every rank
    if kind == "gridworld":
        a = torch.full((2, n_global), 4, dtype=torch.int8, device=device)
        jur = torch.randint(0, 2, (n_global,), device=device, generator=gen)
        pos = torch.randint(0, 4, (n_global,), device=device, generator=gen).to(torch.int8)
        a[0] = torch.where(jur == 0, pos, a[0])
        a[1] = torch.where(jur == 1, pos, a[1])
        return a
    return torch.randint(0, n_actions, (n_cells, n_global), dtype=torch.int8, device=device, generator=gen)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    device = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(device)
    dist.init_process_group("nccl", device_id=device)
    n_global, steps = 200_000 + 48, 9
    report = {}
    cases = [("cellular", B.CellularVectorEnv, dict(kind="cellular", stochastic=True, rng_episodic=False, max_episode_steps=4)),
             ("gridworld", B.CellularVectorEnv, dict(kind="gridworld", max_episode_steps=5, dispersal_prob=0.05)),
             ("packed16x4", B.PackedCellularVectorEnv, dict(n_cells=16, n_states=4, stochastic=True, max_episode_steps=4))]
    for name, cls, kw in cases:
        off, cnt = shard_range(n_global, rank, world)
        env = cls(num_envs=cnt, env_id_offset=off, env_seed=5, device=device, emit_side_effects=False, **kw)
        whole = cls(num_envs=n_global, env_seed=5, device=device, emit_side_effects=False, **kw) if rank == 0 else None
        for t in range(steps):
            a = actions_for(kw.get("kind", "cellular"), env.n_cells, env.n_actions, n_global, t, device)
            env.step_device(a[:, off:off + cnt].contiguous())
            if whole is not None:
                whole.step_device(a)
        mine = torch.cat([env.tabular_state().to(torch.int64), env._reward[:cnt].view(torch.int32).to(torch.int64),
                          env.time_step.to(torch.int64)])
        sizes = [shard_range(n_global, r, world)[1] for r in range(world)]
        bufs = [torch.zeros(3 * c, dtype=torch.int64, device=device) for c in sizes]
        if len(set(sizes)) == 1:
            dist.all_gather(bufs, mine)
        else:                                   # ragged shards: pad to the longest
            m = 3 * max(sizes)
            padded = [torch.zeros(m, dtype=torch.int64, device=device) for _ in sizes]
            dist.all_gather(padded, torch.nn.functional.pad(mine, (0, m - mine.numel())))
            bufs = [p[:3 * c] for p, c in zip(padded, sizes)]
        totals = StatsReducer().start(env._stats).result()
        if rank == 0:
            ok = True
            for r, (buf, c) in enumerate(zip(bufs, sizes)):
                o = shard_range(n_global, r, world)[0]
                ok &= bool((buf[:c] == whole.tabular_state()[o:o + c].to(torch.int64)).all())
                ok &= bool((buf[c:2 * c] == whole._reward[o:o + c].view(torch.int32).to(torch.int64)).all())
                ok &= bool((buf[2 * c:] == whole.time_step[o:o + c].to(torch.int64)).all())
            report[name] = {"shards_equal_whole_batch": ok, "stats_equal": totals == whole.stats(),
                            "env_steps": totals["env_steps"], "expected": n_global * steps}
    if rank == 0:
        print("MULTI_GPU_REPORT " + json.dumps(report), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
