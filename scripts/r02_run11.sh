# round 2, GPU call 11: 32-bit element indexing in all step kernels: tests, bench of every config, shape sweep
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --no-header -rf --timeout 900 ) > gpurun_out/r02_tests11.log 2>&1
tail -5 gpurun_out/r02_tests11.log
show='import json,sys; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d["value"]/1e9,2), "G/s", round(d["ms_per_step"]*1e3,2), "us frac", round(d["roofline"]["frac"],4), "graph", (d.get("cuda_graph") or {}).get("value"), "packed", (d.get("packed") or {}).get("value"))'
for w in cfg4 cfg5 cfg3 cfg2; do
  for i in 1 2; do python bench.py --workload $w --steps 2000 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "$w"; done
done
python scripts/shape_sweep.py 2>&1 | cut -c1-150
