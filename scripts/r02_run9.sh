# round 2, GPU call 9: smoke, full gpu suite, captures for profiles/, final bench lines
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python -m pytest tests -m gpu -q --no-header -rf --timeout 900 ) > gpurun_out/r02_tests9.log 2>&1
tail -5 gpurun_out/r02_tests9.log
bash scripts/capture_profiles_r02.sh > gpurun_out/capture.log 2>&1; tail -2 gpurun_out/capture.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-160 gpurun_out/r02_bench_reference.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_driverlike.json 2> gpurun_out/r02_bench_driverlike.err; tail -c 400 gpurun_out/r02_bench_driverlike.json
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 400 gpurun_out/r02_bench_default.json
show='import json,sys; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d["value"]/1e9,2), "G/s", round(d["ms_per_step"]*1e3,2), "us frac", round(d["roofline"]["frac"],4))'
for i in 1 2 3; do python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "cfg4 20 steps"; done
