"""Multi-GPU parity ON HARDWARE (needs >= 2 GPUs; skipped on a 1-GPU box): one rank per GPU under torchrun
with NCCL.  (1) every rank's shard equals the corresponding slice of the same batch stepped whole on one GPU --
results do not depend on which GPU owns an env; the all-reduced statistics equal the single-GPU totals.
(2) bench.py --gpus 2 prints its line with `shard_check: ok`, config 5 under `workloads`, and consistent
episode statistics."""
import json
import os
import socket
import subprocess
import sys

import pytest

from conftest import REPO

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(n, script, *args, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), script, *args]
    return subprocess.run(cmd, cwd=REPO, capture_output=True, text=True, timeout=timeout)


def _n_gpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def test_shards_on_several_gpus_equal_the_whole_batch():
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    n = 4 if _n_gpus() >= 4 else 2
    out = _torchrun(n, os.path.join(REPO, "tests", "_multi_gpu_worker.py"))
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("MULTI_GPU_REPORT ")][-1]
    report = json.loads(line[len("MULTI_GPU_REPORT "):])
    assert set(report) == {"cellular", "gridworld", "packed16x4"}
    for name, r in report.items():
        assert r["shards_equal_whole_batch"] and r["stats_equal"] and r["env_steps"] == r["expected"], (name, r)


def test_bench_line_on_two_gpus():
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    out = _torchrun(2, "bench.py", "--gpus", "2", "--steps", "40", "--warmup", "3")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1, out.stdout[-800:]
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["shard_check"] == "ok" and d["episode_stats"]["consistent"]
    assert d["gpu_launches"] == 2 * 40 and d["config"]["global_envs"] == 2 * d["config"]["envs_per_gpu"]
    w = d["workloads"]["cfg5"]
    assert w["kernels_per_step"] == 2 and w["episode_stats_consistent"] and w["value"] > 0
    assert d["e2e"]["pcie_measured"]["all_ranks_concurrent_sum"]["d2h_gbs"] > 0
