# round 2, GPU call 43: ncu capture of the config-2 many-step kernel at its final launch geometry (256 blocks of 64 threads)
mkdir -p gpurun_out
B="python bench.py --steps 128 --warmup 64 --no-extra --no-cpu-baseline --no-side"
$B --workload cfg2 > gpurun_out/plain43_cfg2.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:cell_pair_many_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_cfg2_many $B --workload cfg2 > gpurun_out/ncu_43.log 2>&1
ls -la gpurun_out/r02_prof_cfg2_many.ncu-rep
