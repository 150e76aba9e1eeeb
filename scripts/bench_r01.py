#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched gym-cellular step on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4|cfg2|cfg3|cfg5] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (env.step for every env of the batch) over one batch of
synthetic random actions.  Rank 0 prints ONE JSON line (contract in the task statement):
  value      whole-job env-steps/s with actions already resident in HBM (device path, one kernel
             launch per step per GPU), timed with CUDA events, max over ranks
  e2e        the same metric through the host-facing call (numpy in, numpy out: gc_step_host does
             H2D of the actions, the kernel and D2H of observation/reward/index/flags every step)
  roofline   algorithmic HBM bytes per launch / measured launch time, against MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle port (oracle/gc_oracle.c) on all host cores, bounded sample

Workloads (BASELINE.json configs; per-GPU batch fixed => weak scaling):
  cfg4 (default)  polarisation scaled to 16 cells x 4 levels, 2^24 envs per GPU (68 B/env-step,
                  1.14 GB per step: larger than the 126 MB L2, genuinely HBM-bound)
  cfg2            default polarisation env (3 cells x 3 levels), 65,536 envs (1.9 MB: L2-resident)
  cfg3            grid world, 2^20 envs, stochastic dispersal + fused auto-reset (27 MB: L2-resident)
  cfg5            mixed: per GPU 4M stochastic polarisation + 4M grid world envs (64M at 8 GPUs)
The default line also carries cfg2/cfg3/cfg5 results under "workloads" when run on 1 GPU.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

WORKLOADS = {
    # name: description, list of sub-batches (kind, fraction, kwargs), bytes per env-step
    "cfg4": dict(desc="polarisation 16 cells x 4 levels, 2^24 envs per GPU, deterministic (BASELINE config 4)",
                 n_envs=1 << 24, parts=[("cellular", 1.0, dict(n_cells=16, n_states=4))], l2_resident=False),
    "cfg2": dict(desc="polarisation 3 cells x 3 levels, 65,536 envs (BASELINE config 2)",
                 n_envs=1 << 16, parts=[("cellular", 1.0, dict(n_cells=3, n_states=3))], l2_resident=True),
    "cfg3": dict(desc="grid world, 2^20 envs, stochastic dispersal, auto-reset every 128 steps (BASELINE config 3)",
                 n_envs=1 << 20, parts=[("gridworld", 1.0, dict(max_episode_steps=128))], l2_resident=True),
    "cfg5": dict(desc="mixed sweep: half stochastic polarisation (Cells3ResetVDeadlock), half grid world, "
                      "2^23 envs per GPU, auto-reset every 128 steps (BASELINE config 5 = 2^26 envs on 8 GPUs)",
                 n_envs=1 << 23,
                 parts=[("cellular", 0.5, dict(n_cells=3, n_states=3, stochastic=True, max_episode_steps=128)),
                        ("gridworld", 0.5, dict(max_episode_steps=128))], l2_resident=False),
}
RING = 8            # pre-generated action buffers per sub-batch


def bytes_per_env_step(kind, n_cells):
    """Algorithmic HBM bytes of one env-step: state r/w + action r (3C), t r/w (8),
    reward (4), index (4), terminated/truncated/unsafe/count (4)."""
    return 3 * n_cells + 20


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons while the timed region runs (NVML, else nvidia-smi)."""

    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4, "hw_power_brake": 0x80}
    # NVML queries go through the driver and slow concurrent kernel launches down: polled every 5 ms they
    # doubled the per-step time of the launch-bound workloads (15 us against 6-7 us per step)
    PERIOD_S = 0.05

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _poll_once(self):
        if self.nv is not None:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            try:
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for name, bit in {**self.BAD, **self.NOTE}.items():
                if mask & bit:
                    self.reasons.add(name)
        else:
            out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm,"
                                  "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                                  "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
            f = [x.strip() for x in out.split(",")]
            if len(f) >= 6:
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._poll_once()
            except Exception:
                pass
            self._stop.wait(self.PERIOD_S)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        try:
            self._poll_once()           # at least one sample while the last kernels are in flight
        except Exception:
            pass
        self._stop.set()
        self._thread.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------
HOST_CHUNK_ENVS = 1 << 20


def build_batches(workload, device, rank, n_override=None, env_scale=1.0):
    """Creates the vector envs of one rank and their pre-generated device action rings."""
    import torch
    from gym_cellular_b200 import CellularVectorEnv
    w = WORKLOADS[workload]
    n_total = int(n_override or w["n_envs"])
    batches, offset = [], rank * n_total
    gen = torch.Generator(device=device).manual_seed(1234 + rank)
    for kind, frac, kw in w["parts"]:
        n = int(n_total * frac) // 16 * 16
        env = CellularVectorEnv(kind=kind, num_envs=n, device=device, env_seed=0, env_id_offset=offset,
                                emit_side_effects=False, collect_stats=True, host_chunk_envs=HOST_CHUNK_ENVS, **kw)
        offset += n
        ring = []
        for _ in range(RING):
            if kind == "gridworld":
                # the reference sampler's distribution: exactly one jurisdiction names a position (grid_world.py:191-195)
                a = torch.full((2, env.ld), 4, dtype=torch.int8, device=device)
                jur = torch.randint(0, 2, (env.ld,), device=device, generator=gen)
                pos = torch.randint(0, 4, (env.ld,), device=device, generator=gen).to(torch.int8)
                a[0] = torch.where(jur == 0, pos, a[0])
                a[1] = torch.where(jur == 1, pos, a[1])
            else:
                a = torch.randint(0, env.n_actions, (env.n_cells, env.ld), dtype=torch.int8, device=device, generator=gen)
            ring.append(a)
        batches.append(dict(env=env, ring=ring, kind=kind, n=n, bytes=bytes_per_env_step(kind, env.n_cells)))
    return batches



def time_device_path(batches, steps, warmup, dist, device, sampler_index):
    """Device path.  Independent sub-batches (the mixed config 5) step on one stream each, so that
    the tail of one kernel overlaps the head of the other; everything is bracketed by events on the
    main stream, which the side streams are ordered against."""
    import torch
    main = torch.cuda.current_stream(device)
    streams = [main] if len(batches) == 1 else [torch.cuda.Stream(device=device) for _ in batches]
    # one device-path step = one pre-bound ctypes call = one kernel, on the sub-batch's own stream
    calls = [[b["env"].bind_step(a, stream=s) for a in b["ring"]] for b, s in zip(batches, streams)]

    def run(lo, hi):
        for s in streams:
            if s is not main:
                s.wait_stream(main)
        for i in range(lo, hi):
            for c in calls:
                c[i % RING]()
        for s in streams:
            if s is not main:
                main.wait_stream(s)

    run(0, warmup)
    torch.cuda.synchronize(device)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(device)
    launches0 = sum(b["env"].launch_count for b in batches)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(sampler_index) as clk:
        start.record(main)
        run(warmup, warmup + steps)
        stop.record(main)
        stop.synchronize()
    ms = start.elapsed_time(stop)
    launches = sum(b["env"].launch_count for b in batches) - launches0
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    torch.cuda.synchronize(device)
    return ms, launches, clk.summary()


def time_graph_path(batches, steps, device):
    """Launch-bound batch sizes: RING captured steps per CUDA graph, replayed steps/RING times."""
    import torch
    graphs = [b["env"].capture_steps(b["ring"]) for b in batches]
    reps = max(1, steps // RING)
    for g in graphs:
        g.replay()
    torch.cuda.synchronize(device)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(reps):
        for g in graphs:
            g.replay()
    stop.record()
    stop.synchronize()
    return start.elapsed_time(stop), reps * RING


def measure_pcie(device, nbytes=1 << 28):
    """Pinned-memory copy bandwidth of this box (GB/s), for reading the e2e figure: the host path
    moves (C) bytes in and (C + 12) bytes out per env-step and is bound by these two numbers."""
    import torch
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=device)
    out = {}
    for name, (dst, src) in (("h2d_gbs", (d, h)), ("d2h_gbs", (h, d))):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize(device)
        out[name] = round(3 * nbytes / (time.perf_counter() - t0) / 1e9, 1)
    return out


def time_host_path(batches, steps, warmup, dist, device):
    """End to end through the host-facing call: pinned numpy actions in, numpy results out."""
    import torch
    host_rings = []
    for b in batches:
        ring = [b["ring"][i].cpu().pin_memory() for i in range(2)]
        host_rings.append([(t, t.numpy()) for t in ring])
    h2d = sum(b["env"].host_bytes_per_env_step[0] * b["n"] for b in batches)
    d2h = sum(b["env"].host_bytes_per_env_step[1] * b["n"] for b in batches)

    def one(i):
        sink = 0.0
        for b, hr in zip(batches, host_rings):
            obs, rew, term, trunc, info = b["env"].step(hr[i % 2][1][:, :b["n"]])
            sink += float(rew[0])
        return sink
    for i in range(warmup):
        one(i)
    torch.cuda.synchronize(device)
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        one(i)
    torch.cuda.synchronize(device)
    el = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([el], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        el = float(t.item())
    return el, h2d, d2h


# ------------------------------------------------------------------------------------------------
def oracle_envs(workload, n_sample):
    """The CPU oracle configured like `workload` (bench cpu_baseline / --impl reference legs only)."""
    import numpy as np
    from oracle import oracle as O
    out = []
    rng = np.random.default_rng(1234)
    for kind, frac, kw in WORKLOADS[workload]["parts"]:
        n = max(16, int(n_sample * frac))
        if kind == "gridworld":
            env = O.OracleEnv(kind="gridworld", n_envs=n, seed=0, max_episode_steps=kw.get("max_episode_steps", 0))
            a = np.full((2, n), 4, np.int8)
            a[rng.integers(0, 2, n), np.arange(n)] = rng.integers(0, 4, n)
        else:
            st = kw.get("stochastic", False)
            env = O.OracleEnv(n_envs=n, n_cells=kw["n_cells"], n_states=kw["n_states"], noise=st, rng_episodic=True,
                              reward="nonlinear_rp" if st else "right_polarizing", seed=0,
                              max_episode_steps=kw.get("max_episode_steps", 0))
            a = rng.integers(0, kw["n_states"], size=(kw["n_cells"], n)).astype(np.int8)
        out.append((env, a, n))
    return out


def time_oracle(workload, n_sample, steps, warmup, threads):
    envs = oracle_envs(workload, n_sample)
    for _ in range(warmup):
        for env, a, n in envs:
            env.step_parallel(a, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        for env, a, n in envs:
            env.step_parallel(a, threads)
    el = time.perf_counter() - t0
    return sum(n for _, _, n in envs) * steps / el, el, sum(n for _, _, n in envs)


def python_port_rate(workload):
    """What the reference's own execution model (one Python object per env) costs on this box: the
    pure-Python port oracle/pyport.py in a fresh process (fork pool over all cores).  Context only."""
    part = next((kw for kind, _, kw in WORKLOADS[workload]["parts"] if kind == "cellular"), None)
    if part is None:
        return None
    try:
        out = subprocess.run([sys.executable, "-m", "oracle.pyport", "--envs", "64", "--steps", "150", "--cells",
                              str(part["n_cells"]), "--levels", str(part["n_states"])], cwd=REPO, capture_output=True,
                             text=True, timeout=120).stdout.strip().splitlines()[-1]
        return json.loads(out)
    except Exception as exc:
        return {"error": type(exc).__name__}


def cpu_baseline(workload, budget_s=12.0):
    threads = os.cpu_count() or 1
    rate, _, _ = time_oracle(workload, 1 << 16, 2, 1, threads)          # calibration
    n_sample = int(min(1 << 22, max(1 << 14, rate * budget_s / 8)))
    steps = 24
    rate, el, n = time_oracle(workload, n_sample, steps, 1, threads)
    return {"value": rate, "unit": "env-steps/s", "cores": threads, "kind": "port",
            "sample": f"C oracle (oracle/gc_oracle.c), {n} envs x {steps} steps of {workload}, {threads} threads, "
                      f"{el:.1f} s wall = {el * threads:.0f} core-seconds",
            "python_port": python_port_rate(workload)}


def run_reference_arm(args, workload):
    """--impl reference: the CPU implementation on the host cores (the reference is pure Python and
    cannot travel to the GPU box; the oracle port stands in)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rate, _, _ = time_oracle(workload, 1 << 16, 2, 1, threads)
    total = args.steps + args.warmup
    n_sample = int(min(1 << 22, max(1 << 12, rate * 60.0 / max(total, 1))))
    rate, el, n = time_oracle(workload, n_sample, args.steps, args.warmup, threads)
    sample = f"C oracle (oracle/gc_oracle.c), {n} envs per step of {workload}, {threads} threads"
    line = {"impl": "reference", "metric": "env-steps/sec", "value": rate, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8/f64", "data": "synthetic",
            "config": {"workload": workload, "description": WORKLOADS[workload]["desc"], "sample_envs": n},
            "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def bind_near_gpu(device):
    """Pin this rank's host threads (and hence its pinned staging buffers, by first touch) to the CPUs
    of the GPU's NUMA node; matters for the host path when 8 ranks share the host.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device.index)).busId
        bdf = bdf.decode() if isinstance(bdf, bytes) else bdf
        bdf = bdf.lower()
        if len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]
        path = f"/sys/bus/pci/devices/{bdf}/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"{len(allowed)} cpus of the GPU's NUMA node"
        return "no local cpu allowed"
    except Exception as exc:           # pragma: no cover - depends on the box
        return f"unbound ({type(exc).__name__})"


def measured_peak():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload):
    """dram read+write bytes per launch from the committed ncu capture, if any (profiles/traffic.json)."""
    path = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get(workload)
        except Exception:
            return None
    return None


def bench_workload(workload, steps, warmup, dist, device, world, rank, e2e_steps, side=True):
    """`side` = also time the CUDA-graph and fused-rollout variants (single process only)."""
    import torch
    batches = build_batches(workload, device, rank)
    n_rank = sum(b["n"] for b in batches)
    ms, launches, clocks = time_device_path(batches, steps, warmup, dist, device, device.index)
    if not side:
        dist_for_side = True          # any non-None value skips the two side measurements below
    else:
        dist_for_side = dist
    graph_res = None
    if WORKLOADS[workload]["l2_resident"] and dist_for_side is None:
        g_ms, g_steps = time_graph_path(batches, steps, device)
        graph_res = {"value": n_rank * g_steps / (g_ms * 1e-3), "ms_per_step": g_ms / g_steps,
                     "steps_per_graph": RING}
    ro_res = None
    if dist_for_side is None:              # K-step fused rollout (random actions generated in the kernel)
        K = 64
        main = torch.cuda.current_stream(device)
        ro_streams = [main] if len(batches) == 1 else [torch.cuda.Stream(device=device) for _ in batches]

        def rollouts(reps):                # independent sub-batches on one stream each, as in the step path
            for s in ro_streams:
                if s is not main:
                    s.wait_stream(main)
            for _ in range(reps):
                for b, s in zip(batches, ro_streams):
                    with torch.cuda.stream(s):
                        b["env"].rollout(K)
            for s in ro_streams:
                if s is not main:
                    main.wait_stream(s)
        rollouts(1)
        torch.cuda.synchronize(device)
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 4
        r0.record(main)
        rollouts(reps)
        r1.record(main)
        r1.synchronize()
        ro_res = {"value": n_rank * K * reps / (r0.elapsed_time(r1) * 1e-3), "steps_per_launch": K,
                  "note": "fused rollout: state in registers, actions generated in-kernel (not per-step step())"}
    el_host, h2d, d2h = time_host_path(batches, e2e_steps, 2, dist, device)
    # the only collective of the path: episode statistics, all-reduced once per iteration (NCCL, side stream)
    from gym_cellular_b200.distributed import StatsReducer
    totals = StatsReducer().start(torch.stack([b["env"]._stats for b in batches]).sum(0)).result()
    alg_bytes = sum(b["bytes"] * b["n"] for b in batches)           # per step, per rank
    peak, peak_src = measured_peak()
    step_s = ms * 1e-3 / steps
    res = {
        "value": world * n_rank * steps / (ms * 1e-3),
        "ms_per_step": ms / steps,
        "envs_per_gpu": n_rank,
        "gpu_launches": launches * world,
        "clocks": clocks,
        "e2e": {"value": world * n_rank * e2e_steps / el_host, "unit": "env-steps/s",
                "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world, "steps": e2e_steps},
        "roofline": {"bound": "hbm", "achieved": alg_bytes / step_s / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg_bytes / step_s / 1e9 / peak, "traffic": ncu_traffic(workload),
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes // len(batches),
                     "bytes_per_env_step": alg_bytes / n_rank, "kernels_per_step": len(batches),
                     "l2_resident": WORKLOADS[workload]["l2_resident"]},
        "episode_stats": totals,
        "cuda_graph": graph_res,
        "fused_rollout": ro_res,
    }
    for b in batches:
        b["env"].close()
    del batches
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg2/cfg3/cfg5 side measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side", action="store_true",
                    help="skip the CUDA-graph and fused-rollout side measurements (profiling runs: launch lists)")
    ap.add_argument("--host-chunk", type=int, default=None, help="envs per chunk of the host (e2e) path")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.host_chunk:
        global HOST_CHUNK_ENVS
        HOST_CHUNK_ENVS = args.host_chunk

    if args.impl == "reference":
        run_reference_arm(args, args.workload)
        return

    # stdout carries the one JSON line and nothing else: libraries that write to file descriptor 1 behind
    # Python's back (NCCL prints its version banner there, and its log at NCCL_DEBUG=INFO) are sent to
    # stderr by pointing descriptor 1 at it; Python's own sys.stdout keeps the original descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    affinity = bind_near_gpu(device) if world > 1 else "single rank: unbound"
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    e2e_steps = max(3, min(args.steps, 10))
    main_res = bench_workload(args.workload, args.steps, args.warmup, dist, device, world, rank, e2e_steps,
                              side=not args.no_side)
    extra = {}
    if world == 1 and not args.no_extra:
        for w in ("cfg2", "cfg3", "cfg5"):
            if w != args.workload:
                r = bench_workload(w, args.steps, args.warmup, None, device, 1, 0, e2e_steps)
                extra[w] = {"description": WORKLOADS[w]["desc"], "value": r["value"], "ms_per_step": r["ms_per_step"],
                            "roofline_frac": r["roofline"]["frac"], "achieved_gbs": r["roofline"]["achieved"],
                            "l2_resident": WORKLOADS[w]["l2_resident"], "e2e": r["e2e"]["value"],
                            "kernels_per_step": r["roofline"]["kernels_per_step"], "cuda_graph": r["cuda_graph"], "fused_rollout": r["fused_rollout"]}
    if rank == 0:
        line = {
            "metric": "env-steps/sec", "value": main_res["value"], "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8/int32 state, f32 reward",
            "data": "synthetic",
            "config": {"workload": args.workload, "description": WORKLOADS[args.workload]["desc"],
                       "envs_per_gpu": main_res["envs_per_gpu"], "global_envs": main_res["envs_per_gpu"] * world,
                       "parallelism": f"env-sharded x{world}, NCCL all-reduce of episode statistics once per iteration",
                       "l2": "per-step working set larger than L2" if not WORKLOADS[args.workload]["l2_resident"]
                             else "working set is L2-resident (launch-bound, not an HBM measurement)",
                       "actions": f"ring of {RING} pre-generated device buffers, uniform random",
                       "host_affinity": affinity},
            "e2e": {**main_res["e2e"], "pcie_measured": measure_pcie(device)},
            "gpu_launches": main_res["gpu_launches"], "clocks": main_res["clocks"],
            "roofline": main_res["roofline"], "episode_stats": main_res["episode_stats"],
            "fused_rollout": main_res["fused_rollout"], "cuda_graph": main_res["cuda_graph"],
        }
        if extra:
            line["workloads"] = extra
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
