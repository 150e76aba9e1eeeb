# round 2, GPU call 12: grid-world kernel diet (time limit, reward conversion): tests + configs 3 and 5
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --no-header -rf --timeout 900 ) > gpurun_out/r02_tests12.log 2>&1
tail -4 gpurun_out/r02_tests12.log
show='import json,sys; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d["value"]/1e9,2), "G/s", round(d["ms_per_step"]*1e3,2), "us frac", round(d["roofline"]["frac"],4), "graph", (d.get("cuda_graph") or {}).get("value"))'
for w in cfg5 cfg3; do
  for i in 1 2 3; do python bench.py --workload $w --steps 2000 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "$w"; done
done
python scripts/shape_sweep.py 2>&1 | grep gridworld | cut -c1-150
