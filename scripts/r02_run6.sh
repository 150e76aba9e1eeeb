# round 2, GPU call 6: tests (pair8 kernel, graph-backed gc_step_many), short-run comparison with the round-1 bench, sweeps
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --no-header -rf --timeout 900 ) > gpurun_out/r02_tests6.log 2>&1
tail -6 gpurun_out/r02_tests6.log
show='import json,sys; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d["value"]/1e9,2), "G/s", round(d["ms_per_step"]*1e3,2), "us frac", round(d["roofline"]["frac"],4), "host_us", d.get("host_us_per_launch"), "graph", (d.get("cuda_graph") or {}).get("value"))'
for i in 1 2 3; do python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "r02 bench 20 steps"; done
for i in 1 2 3; do python scripts/bench_r01.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "r01 bench 20 steps"; done
python bench.py --steps 2000 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "r02 bench 2000 steps"
python scripts/bench_r01.py --steps 2000 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "r01 bench 2000 steps"
for w in cfg2 cfg3 cfg5; do
  python bench.py --workload $w --steps 2000 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "r02 $w 2000 steps"
  python bench.py --workload $w --steps 500 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "r02 $w 500 steps"
  GC_B200_STEP_MANY_GRAPH=0 python bench.py --workload $w --steps 2000 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "r02 $w 2000 steps, no graph in step_many"
done
python scripts/shape_sweep.py > gpurun_out/r02e_shapes.txt 2>&1; cat gpurun_out/r02e_shapes.txt
