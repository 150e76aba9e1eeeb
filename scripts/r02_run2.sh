# round 2, GPU call 2: full gpu test suite (all failures), packed kernel after the instruction diet, bench.py end to end
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --no-header -rf --timeout 900 ) > gpurun_out/r02_tests2.log 2>&1
tail -15 gpurun_out/r02_tests2.log
P="python scripts/packed_bench.py"
$P --no-host > gpurun_out/r02b_packed_cfg4.json 2>&1; cat gpurun_out/r02b_packed_cfg4.json
$P --envs 8388608 --cells 3 --levels 3 --no-host --stochastic > gpurun_out/r02b_packed_c3s.json 2>&1; cat gpurun_out/r02b_packed_c3s.json
$P --envs 65536 --cells 3 --levels 3 --no-host --steps 3000 > gpurun_out/r02b_packed_cfg2.json 2>&1; cat gpurun_out/r02b_packed_cfg2.json
( time python bench.py ) > gpurun_out/r02b_bench_default.json 2> gpurun_out/r02b_bench_default.err
tail -c 3000 gpurun_out/r02b_bench_default.json; tail -5 gpurun_out/r02b_bench_default.err
ncu --set full --clock-control none --import-source on -k regex:cell_packed_kernel -s 12 -c 2 -f -o gpurun_out/r02b_prof_packed $P --no-host --no-int8 --steps 20 > gpurun_out/ncu_packed2.log 2>&1
