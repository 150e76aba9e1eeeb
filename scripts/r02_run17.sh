# round 2, GPU call 17: short-run check of the gated start, example test
show='import json,sys; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d["value"]/1e9,2), "G/s", round(d["ms_per_step"]*1e3,2), "us frac", round(d["roofline"]["frac"],4), "packed", round(d["packed"]["value"]/1e9,1) if d.get("packed") else None)'
for i in 1 2 3; do python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "cfg4 20 steps, gated start"; done
python bench.py --steps 2000 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "cfg4 2000 steps"
for w in cfg5 cfg3 cfg2; do python bench.py --workload $w --steps 20 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "$w 20 steps"; done
python -m pytest tests/test_gpu_packed.py tests/test_bench_contract.py -m gpu -q --no-header -rf -k "example or native_arm" 2>&1 | tail -3
