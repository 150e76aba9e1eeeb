# round 2, GPU call 8: pair-table replication A/B on the int8 config-4 kernel, max log2 error, final captures and bench lines
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -s -k "log2_reward" 2>&1 | grep -i "max relative\|passed\|failed"
show='import json,sys; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d["value"]/1e9,2), "G/s", round(d["ms_per_step"]*1e3,2), "us frac", round(d["roofline"]["frac"],4))'
for i in 1 2; do python bench.py --steps 2000 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "cfg4 int8, plain pair table"; done
python scripts/shape_sweep.py --cells 2>&1 | grep "'stochastic': False" | grep "n_cells': 1[2-6]"
GC_NVCC_EXTRA="-DGC_PAIR_REP_LOG2=4" python -m gym_cellular_b200.build --force > /dev/null 2>&1
for i in 1 2; do python bench.py --steps 2000 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "cfg4 int8, pair table replicated 16x"; done
python scripts/shape_sweep.py --cells 2>&1 | grep "'stochastic': False" | grep "n_cells': 1[2-6]"
GC_NVCC_EXTRA="-DGC_PAIR_REP_LOG2=3" python -m gym_cellular_b200.build --force > /dev/null 2>&1
for i in 1 2; do python bench.py --steps 2000 --no-extra --no-cpu-baseline 2>/dev/null | python -c "$show" "cfg4 int8, pair table replicated 8x"; done
python -m gym_cellular_b200.build --force > /dev/null 2>&1
