# round 2, GPU call 29: the committed library after the wide-noise series: full suite, smoke, bench lines, sweep, captures of the wide stochastic kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --no-header -rf --timeout 900 -x > gpurun_out/r02_tests29.log 2>&1; tail -3 gpurun_out/r02_tests29.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke29.log 2>&1; tail -2 gpurun_out/r02_smoke29.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -c 200 gpurun_out/r02_bench_reference.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_driverlike.json 2> gpurun_out/r02_bench_driverlike.err; tail -c 150 gpurun_out/r02_bench_driverlike.json
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 150 gpurun_out/r02_bench_default.json
python scripts/shape_sweep.py > gpurun_out/r02_sweep29.log 2>&1
python scripts/shape_sweep.py --only 6 --packed >> gpurun_out/r02_sweep29.log 2>&1
cat gpurun_out/r02_sweep29.log
ncu --set full --clock-control none --import-source on -k regex:cell_pair_kernel -s 10 -c 1 -f -o gpurun_out/r02_prof_16x4_noise python scripts/shape_sweep.py --only 6 > gpurun_out/ncu_29a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cell_packed_kernel -s 10 -c 1 -f -o gpurun_out/r02_prof_16x4_noise_packed python scripts/shape_sweep.py --only 6 --packed > gpurun_out/ncu_29b.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
