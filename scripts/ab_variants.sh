# A/B of compile-time kernel variants on ONE box (rebuilds on the GPU box with nvcc)
set -e
for v in "-DGC_PAIR_PREFETCH_WIDE=0 -DGC_PAIR_MINB=4" "-DGC_PAIR_PREFETCH_WIDE=1 -DGC_PAIR_MINB=4" "-DGC_PAIR_PREFETCH_WIDE=1 -DGC_PAIR_MINB=3" "-DGC_PAIR_PREFETCH_WIDE=0 -DGC_PAIR_MINB=3"; do
  GC_NVCC_EXTRA="$v" python -m gym_cellular_b200.build --force > /dev/null 2>&1
  for w in cfg4; do
    for rep in 1 2; do
    python bench.py --workload $w --steps 1000 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('variant[$v] $w', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step']*1e3,1), 'us frac', round(d['roofline']['frac'],3))"
    done
  done
done
python -m gym_cellular_b200.build --force > /dev/null 2>&1
