# round 2, GPU call 34: full suite and the bench lines with gc_step_many in one launch for small shards
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --no-header -rf --timeout 900 -x > gpurun_out/r02_tests34.log 2>&1; tail -3 gpurun_out/r02_tests34.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke34.log 2>&1; tail -1 gpurun_out/r02_smoke34.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_driverlike.json 2> gpurun_out/r02_bench_driverlike.err; tail -c 150 gpurun_out/r02_bench_driverlike.json; echo
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 150 gpurun_out/r02_bench_default.json; echo
