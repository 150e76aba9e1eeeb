# round 2, GPU call 15 (final 1-GPU): smoke, gpu suite, captures for profiles/, bench lines
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python -m pytest tests -m gpu -q --no-header -rf --timeout 900 ) > gpurun_out/r02_tests15.log 2>&1
grep -E "passed|failed" gpurun_out/r02_tests15.log
bash scripts/capture_profiles_r02.sh > gpurun_out/capture.log 2>&1; tail -1 gpurun_out/capture.log; du -sh gpurun_out
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-160 gpurun_out/r02_bench_reference.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_driverlike.json 2> gpurun_out/r02_bench_driverlike.err; tail -c 200 gpurun_out/r02_bench_driverlike.json
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 200 gpurun_out/r02_bench_default.json
du -sh gpurun_out
