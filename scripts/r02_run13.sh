# round 2, GPU call 13: 32-step graphs in gc_step_many, lighter grid-world statistics flush: tests + configs 2, 3, 5
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --no-header -rf --timeout 900 ) > gpurun_out/r02_tests13.log 2>&1
grep -E "passed|failed" gpurun_out/r02_tests13.log
show='import json,sys; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d["value"]/1e9,2), "G/s", round(d["ms_per_step"]*1e3,2), "us frac", round(d["roofline"]["frac"],4), "graph", (d.get("cuda_graph") or {}).get("value"))'
for w in cfg3 cfg2 cfg5; do
  for i in 1 2 3; do python bench.py --workload $w --steps 2000 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "$w"; done
done
python bench.py --workload cfg3 --steps 20 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "cfg3 main workload, 20 steps"
