"""N > 1 host logic on CPU (gloo, world_size 2): shard arithmetic, statistics all-reduce, and
world-size invariance of a sharded rollout (stepped by the CPU oracle, since there is no GPU here)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gym_cellular_b200.distributed import StatsReducer, shard_range


def test_shard_range_tiles_the_batch():
    for n in (1, 15, 16, 17, 1000, 65536, (1 << 26) + 5):
        for world in (1, 2, 3, 4, 8):
            pieces = [shard_range(n, r, world) for r in range(world)]
            assert pieces[0][0] == 0 and sum(c for _, c in pieces) == n
            for (o1, c1), (o2, _) in zip(pieces, pieces[1:]):
                assert o1 + c1 == o2 or c1 == 0
            assert all(o % 16 == 0 for o, _ in pieces)
            if n >= 16 * world:
                sizes = [c for _, c in pieces]
                assert max(sizes) - min(sizes) <= 16 + n % 16


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_global, steps, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        off, cnt = shard_range(n_global, rank, world)
        env = O.OracleEnv(n_envs=cnt, noise=True, seed=42, env_id_offset=off, max_episode_steps=5, reward="nonlinear_rp")
        rng = np.random.default_rng(7)
        for _ in range(steps):
            a = rng.integers(0, 3, size=(3, n_global)).astype(np.int8)[:, off:off + cnt]   # same global action stream
            env.step(np.ascontiguousarray(a))
        red = StatsReducer().start(torch.from_numpy(env.stats.copy()))
        totals = red.result()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), state=env.state, off=off, stats=env.stats,
                 totals=np.array([totals["env_steps"], totals["unsafe_steps"], totals["count_sum"],
                                  totals["episodes_truncated"]]))
    finally:
        dist.destroy_process_group()


def test_two_rank_rollout_equals_single_process(tmp_path):
    from oracle import oracle as O
    n_global, steps, world = 1000, 12, 2
    mp.spawn(_worker, args=(world, _free_port(), n_global, steps, str(tmp_path)), nprocs=world, join=True)
    ref = O.OracleEnv(n_envs=n_global, noise=True, seed=42, max_episode_steps=5, reward="nonlinear_rp")
    rng = np.random.default_rng(7)
    for _ in range(steps):
        ref.step(rng.integers(0, 3, size=(3, n_global)).astype(np.int8))
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert (np.concatenate([p["state"] for p in parts], axis=1) == ref.state).all()
    for p in parts:                                   # every rank holds the global totals
        assert (p["totals"] == ref.stats[:4]).all()
    assert sum(int(p["stats"][4]) for p in parts) == ref.stats[4]     # fixed-point reward sum adds exactly


def test_reducer_without_process_group():
    s = torch.tensor([10, 1, 2, 3, 5 << 24, 0, 0, 0])
    assert StatsReducer().start(s).result() == {"env_steps": 10, "unsafe_steps": 1, "count_sum": 2,
                                                 "episodes_truncated": 3, "reward_sum": 5.0}
