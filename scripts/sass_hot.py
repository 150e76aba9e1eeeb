#!/usr/bin/env python
"""Per-instruction executed counts and stall samples of one kernel from an .ncu-rep (source page).

    python scripts/sass_hot.py gpurun_out/r01_prof_cfg5.ncu-rep cell_pair [launch_index]
"""
import csv, io, subprocess, sys

rep, pat = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
blocks = raw.split('"Kernel Name",')[1:]
rows = list(csv.reader(io.StringIO(blocks[which].split("\n", 1)[1])))
hdr = rows[0]
i_src, i_ex, i_smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot_ex = sum(int(r[i_ex]) for r in rows[1:] if len(r) > i_smp)
tot_smp = sum(int(r[i_smp]) for r in rows[1:] if len(r) > i_smp)
print(f"# {blocks[which].splitlines()[0][:100]}  executed={tot_ex} samples={tot_smp}")
for n, r in enumerate(rows[1:]):
    if len(r) <= i_smp:
        continue
    print(f"{n:5d} {int(r[i_ex]):9d} {int(r[i_smp]):6d}  {r[i_src].strip()}")
