# Capture recipe for profiles/: plain runs first, then ncu launch lists and --set full reports.
mkdir -p gpurun_out
set -x
B="python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline --no-side"
$B > gpurun_out/plain.log 2>&1 || exit 1
$B --workload cfg5 > gpurun_out/plain5.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_cfg4.csv $B > gpurun_out/ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_cfg5.csv $B --workload cfg5 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cell_pair_kernel -s 6 -c 2 -f -o gpurun_out/r01_prof_cfg4 $B > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"grid_step_kernel|cell_pair_kernel" -s 12 -c 4 -f -o gpurun_out/r01_prof_cfg5 $B --workload cfg5 > gpurun_out/ncu4.log 2>&1
python bench.py > gpurun_out/r01_bench_default.json 2> gpurun_out/r01_bench_default.err
tail -c 600 gpurun_out/r01_bench_default.json
