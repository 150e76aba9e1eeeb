# register-staged vs TMA bulk-staged kernel for wide deterministic envs, same box
set -e
for st in 1 2; do
  GC_NVCC_EXTRA="-DGC_TMA_STAGES=$st" python -m gym_cellular_b200.build --force > /dev/null 2>&1
  python -m pytest tests/test_gpu_parity.py -x -q -k "config4 or full_size" 2>&1 | tail -1
  for v in 0 1; do
    GC_B200_TMA=$v python bench.py --workload cfg4 --steps 1000 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('stages=$st GC_B200_TMA=$v cfg4', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step']*1e3,1), 'us frac', round(d['roofline']['frac'],3))"
  done
done
python -m gym_cellular_b200.build --force > /dev/null 2>&1
