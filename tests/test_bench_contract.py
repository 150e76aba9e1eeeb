"""bench.py prints ONE JSON line with the keys the driver reads (CPU: the reference arm; GPU: the native arm)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import REPO

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(*args):
    out = subprocess.run([sys.executable, "bench.py", *args], cwd=REPO, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-1500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[-500:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run("--impl", "reference", "--steps", "3", "--warmup", "3", "--workload", "cfg5")
    assert d["impl"] == "reference" and BASE_KEYS <= d.keys()
    assert d["value"] > 0 and d["unit"] == "env-steps/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "cfg5" and d["gpu_launches"] == 0


@pytest.mark.gpu
def test_native_arm_line():
    d = _run("--steps", "30", "--warmup", "3", "--workload", "cfg2", "--no-extra")
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks"} <= d.keys() and "impl" not in d
    # a 65,536-env shard: the 30 bound steps of the gc_step_many call run inside one kernel, and the same loop is
    # reported with one launch per step beside it
    if os.environ.get("GC_B200_STEP_MANY_FUSED", "1")[:1] == "0":
        pytest.skip("the one-launch path is switched off in this environment")
    assert d["gpu_launches"] == 1 and d["steps_per_launch"] == 30 and d["separate_launches"]["gpu_launches"] == 30
    assert 0 < d["separate_launches"]["value"] < d["value"] and d["episode_stats"]["consistent"]
    assert d["n_gpus"] == 1 and d["scaling"] == "weak" and d["data"] == "synthetic"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # (the many-step kernel reads state and t once per launch: 29 - 7 bytes per env-step)
    assert r["bytes_per_env_step"] == 22.0 and r["algorithmic_bytes_per_step"] == 22 * 65536
    assert r["algorithmic_bytes_per_launch"] == 30 * 22 * 65536
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 4 * 65536 and e["d2h_bytes_per_step"] == 13 * 65536 and 0 < e["value"] < d["value"]
    assert d["clocks"]["sm_max_mhz"] and not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
    assert d["cpu_baseline"]["kind"] == "port" and d["config"]["workload"] == "cfg2"
