# round 2, GPU call 35: ncu captures of the many-step kernels (configs 2 and 3 through gc_step_many)
mkdir -p gpurun_out
B="python bench.py --steps 128 --warmup 64 --no-extra --no-cpu-baseline --no-side"
for w in cfg2 cfg3; do $B --workload $w > gpurun_out/plain35_$w.log 2>&1 || exit 1; done
ncu --set full --clock-control none --import-source on -k regex:cell_pair_many_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_cfg2_many $B --workload cfg2 > gpurun_out/ncu_35a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:grid_many_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_cfg3_many $B --workload cfg3 > gpurun_out/ncu_35b.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_cfg3_many.csv $B --workload cfg3 > gpurun_out/ncu_35c.log 2>&1
ls -la gpurun_out/r02_prof_cfg*_many.ncu-rep
