# round 2, GPU call 7: tests, short and long bench runs with the prepared step_many graphs, pair8 sweep
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --no-header -rf --timeout 900 ) > gpurun_out/r02_tests7.log 2>&1
tail -6 gpurun_out/r02_tests7.log
show='import json,sys; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d["value"]/1e9,2), "G/s", round(d["ms_per_step"]*1e3,2), "us frac", round(d["roofline"]["frac"],4), "host_us", d.get("host_us_per_launch"), "graph", (d.get("cuda_graph") or {}).get("value"), "packed", (d.get("packed") or {}).get("value"))'
for i in 1 2 3; do python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "r02 cfg4 20 steps"; done
python scripts/bench_r01.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "r01 cfg4 20 steps"
for w in cfg2 cfg3 cfg5; do
  python bench.py --workload $w --steps 2000 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "r02 $w 2000 steps"
  python bench.py --workload $w --steps 500 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "r02 $w 500 steps"
  python bench.py --workload $w --steps 20 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "r02 $w 20 steps"
done
python scripts/shape_sweep.py 2>&1 | grep "n_states': [58]"
GC_NVCC_EXTRA="-DGC_PAIR8_MINB=3" python -m gym_cellular_b200.build --force > /dev/null 2>&1
echo "pair8 three blocks:"; python scripts/shape_sweep.py 2>&1 | grep "n_states': [58]"
