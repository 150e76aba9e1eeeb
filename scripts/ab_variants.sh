# A/B of compile-time kernel variants on ONE box (rebuilds on the GPU box with nvcc)
set -e
for v in "-DGC_PAIR_MINB_NARROW=4" "-DGC_PAIR_MINB_NARROW=5" "-DGC_PAIR_MINB_NARROW=6"; do
  GC_NVCC_EXTRA="$v" python -m gym_cellular_b200.build --force > /dev/null 2>&1
  for w in cfg5; do
    python bench.py --workload $w --steps 1000 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('variant[$v] $w', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step']*1e3,1), 'us frac', round(d['roofline']['frac'],3))"
  done
done
python -m gym_cellular_b200.build --force > /dev/null 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
