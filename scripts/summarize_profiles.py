#!/usr/bin/env python
"""Turns the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py r01

Reads  gpurun_out/<round>_launches_<workload>.csv   (ncu --metrics gpu__time_duration.sum --csv)
       gpurun_out/<round>_prof_<workload>.ncu-rep   (ncu --set full --import-source on)
Writes profiles/<round>_launches_<workload>.md, profiles/<round>_kernel_<workload>.md and
       profiles/traffic.json (dram read+write bytes per launch of the dominant kernel, for bench.py).
"""
import csv
import glob
import io
import json
import os
import re
import subprocess
import sys
from collections import Counter, defaultdict

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "profiles")
SRC = os.path.join(REPO, "gpurun_out")

RAW_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(value.replace(",", "")) * scale


DEVICE_LAUNCHES = 23      # --steps 20 --warmup 3 of scripts/capture_profiles*.sh


def launches(path, round_, workload):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    per = defaultdict(list)
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        per[re.sub(r"\s+", " ", r[ki])[:110]].append(v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3}.get(r[ui], 1e-3))
    # the step kernels are launched over the whole shard by the device path (`value`) and over 1M-env
    # chunks by the host path (`e2e`): two rows per kernel, split at half the longest launch
    for k in [k for k in per if "_step_kernel" in k or "cell_pair_kernel" in k or "cell_packed_kernel" in k]:
        v = per.pop(k)
        cut = max(v) / 2
        for i, grp in enumerate(([x for x in v if x >= cut], [x for x in v if x < cut])):
            if grp:          # the capture command steps 20 + 3 warm-up times: DEVICE_LAUNCHES whole-shard launches
                tag = " [whole shard: device path]" if i == 0 else " [1M-env chunks: host path]"
                per[k + tag] = per.get(k + tag, []) + grp
    total = sum(sum(v) for v in per.values())
    out = [f"# {round_}: ncu launch list, workload {workload}", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: compare",
           "shares, not absolutes).  Our kernels are the `<unnamed>::*_kernel` rows; `at::` rows are torch's",
           "set-up work (action-ring generation, zero fills) outside the timed region.", "",
           "| kernel | launches | avg us | total us | share |", "|---|---|---|---|---|"]
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k}` | {len(v)} | {sum(v) / len(v):.1f} | {sum(v):.0f} | {100 * sum(v) / total:.1f}% |")
    out += ["", "Reading: the timed region of the device path (`value`) launches nothing but the whole-shard rows",
            "(one kernel of ours per step and sub-batch, `gpu_launches` = steps): the step kernels are 100 % of a step.",
            "`cell_packed_kernel` whole-shard rows are the packed-layout device path (`packed`), its 1M-env chunk rows the",
            "host path (`gc_step_host_packed`, the `e2e` leg); `tick_kernel` advances the device-resident step counter once",
            "per host-path step; `pack_kernel` converts the pre-generated action ring once at set-up; everything `at::` is",
            "torch set-up work outside the timed regions (action-ring generation, zero fills) or the 64-byte statistics",
            "snapshot taken once per 64-step iteration."]
    open(os.path.join(OUT, f"{round_}_launches_{workload}.md"), "w").write("\n".join(out) + "\n")


def kernel_report(rep, round_, workload, traffic):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = [f"# {round_}: ncu --set full, workload {workload}", "",
           f"source: `gpurun_out/{os.path.basename(rep)}` (`ncu --set full --clock-control none --import-source on`)", ""]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        out += [f"## `{name}`", "", "| metric | value | unit |", "|---|---|---|"]
        for k in RAW_KEYS:
            if k in hdr:
                out.append(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
        wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        out += ["", f"DRAM traffic per launch: {(rd + wr) / 1e6:.1f} MB (read {rd / 1e6:.1f} + write {wr / 1e6:.1f}; "
                "writes still resident in the 126 MB L2 at kernel end are not counted by the DRAM counters)", ""]
        traffic.setdefault(workload, {})[name] = rd + wr
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name",
                              "regex:" + re.escape(name.split("(")[0].split("::")[-1].split("<")[0])],
                             capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        if len(srows) > 2 and "Instructions Executed" in srows[1]:
            h = srows[1]
            si, ei = h.index("Source"), h.index("Instructions Executed")
            ops, tot = Counter(), 0
            for sr in srows[2:]:
                try:
                    n = int(sr[ei])
                except (ValueError, IndexError):
                    continue
                m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", sr[si])
                ops[m.group(2) if m else "?"] += n
                tot += n
            out += ["Executed warp instructions by opcode (source page, all launches of this kernel name in the report):", "",
                    "| opcode | warp instr | share |", "|---|---|---|"]
            out += [f"| {op} | {n} | {100 * n / tot:.1f}% |" for op, n in ops.most_common(16)]
            out.append("")
    open(os.path.join(OUT, f"{round_}_kernel_{workload}.md"), "w").write("\n".join(out) + "\n")


def main():
    round_ = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    traffic = {}
    tpath = os.path.join(OUT, "traffic.json")
    old = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for path in sorted(glob.glob(os.path.join(SRC, f"{round_}_launches_*.csv"))):
        launches(path, round_, re.search(r"_launches_(\w+)\.csv", path).group(1))
    for rep in sorted(glob.glob(os.path.join(SRC, f"{round_}_prof_*.ncu-rep"))):
        kernel_report(rep, round_, re.search(r"_prof_(\w+)\.ncu-rep", rep).group(1), traffic)
    flat = dict(old)                      # workloads not re-captured this round keep their entry
    for w, d in traffic.items():
        flat[w] = sum(d.values()) if w.startswith("cfg5") else max(d.values())
        flat[w + "_per_kernel"] = d
        flat[w + "_round"] = round_
    json.dump(flat, open(tpath, "w"), indent=1)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
