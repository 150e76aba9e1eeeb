# round 2, GPU call 18: narrow envs share a Philox block between two envs: tests, config 5, register budget A/B
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --no-header -rf --timeout 900 ) > gpurun_out/r02_tests18.log 2>&1
grep -E "passed|failed" gpurun_out/r02_tests18.log; grep -E "^FAILED" gpurun_out/r02_tests18.log | head
show='import json,sys; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d["value"]/1e9,2), "G/s", round(d["ms_per_step"]*1e3,2), "us frac", round(d["roofline"]["frac"],4), "packed", round(d["packed"]["value"]/1e9,1) if d.get("packed") else None)'
for i in 1 2; do python bench.py --workload cfg5 --steps 2000 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "cfg5, 4 blocks"; done
python scripts/shape_sweep.py 2>&1 | grep "n_cells': [23], 'n_states': 3" | cut -c1-150
GC_NVCC_EXTRA="-DGC_PAIR_MINB_NARROW_RNG=3" python -m gym_cellular_b200.build --force > /dev/null 2>&1
for i in 1 2; do python bench.py --workload cfg5 --steps 2000 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "cfg5, 3 blocks for the narrow Philox kernels"; done
python scripts/shape_sweep.py 2>&1 | grep "n_cells': [23], 'n_states': 3" | cut -c1-150
