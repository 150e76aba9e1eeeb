# (gpurun brings back at most 64 MiB: the cfg5 capture takes two launches of each of its two kernels -- with
# gc_step_many the launch order is 3 + 3 warm-up launches, then 20 + 20 --, cfg2 / cfg3 go without the source page)
# Round-2 capture recipe for profiles/: plain runs first (must exit 0), then ncu launch lists and --set full reports.
mkdir -p gpurun_out
set -x
B="python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline --no-side"
for w in cfg4 cfg5 cfg2 cfg3; do $B --workload $w > gpurun_out/plain_$w.log 2>&1 || exit 1; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_cfg4.csv $B > gpurun_out/ncu_l4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_cfg5.csv $B --workload cfg5 > gpurun_out/ncu_l5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cell_pair_kernel -s 6 -c 2 -f -o gpurun_out/r02_prof_cfg4 $B > gpurun_out/ncu_p4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cell_packed_kernel -s 6 -c 2 -f -o gpurun_out/r02_prof_cfg4_packed $B > gpurun_out/ncu_p4p.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"grid_step_kernel|cell_pair_kernel" -s 4 -c 4 -f -o gpurun_out/r02_prof_cfg5 $B --workload cfg5 > gpurun_out/ncu_p5.log 2>&1
ncu --set full --clock-control none -k regex:cell_pair_kernel -s 6 -c 2 -f -o gpurun_out/r02_prof_cfg2 $B --workload cfg2 > gpurun_out/ncu_p2.log 2>&1
ncu --set full --clock-control none -k regex:grid_step_kernel -s 6 -c 2 -f -o gpurun_out/r02_prof_cfg3 $B --workload cfg3 > gpurun_out/ncu_p3.log 2>&1
set +x
