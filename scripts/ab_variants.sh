# A/B of compile-time kernel variants on ONE box (rebuilds on the GPU box with nvcc)
for v in "" "-DGC_OVERSUB=2 -DGC_GRID_OVERSUB=2" "-DGC_OVERSUB=4 -DGC_GRID_OVERSUB=4" "-DGC_OVERSUB=16 -DGC_GRID_OVERSUB=16" "-DGC_OVERSUB=4" "-DGC_GRID_OVERSUB=4"; do
  GC_NVCC_EXTRA="$v" python -m gym_cellular_b200.build --force > /dev/null 2>&1
  for w in cfg5 cfg4; do
    python bench.py --workload $w --steps 2000 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('variant[$v] $w', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step']*1e3,1), 'us frac', round(d['roofline']['frac'],3))"
  done
done
python -m gym_cellular_b200.build --force > /dev/null 2>&1
