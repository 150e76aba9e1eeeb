# A/B of compile-time kernel variants on ONE box (rebuilds on the GPU box with nvcc)
for v in "-DGC_GRID_THREADS=512 -DGC_GRID_MINB=2" "" "-DGC_GRID_THREADS=512 -DGC_GRID_MINB=2" ""; do
  GC_NVCC_EXTRA="$v" python -m gym_cellular_b200.build --force > /dev/null 2>&1
  for w in cfg3 cfg3; do
  python bench.py --workload cfg3 --steps 3000 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('variant[$v] cfg3', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step']*1e3,1), 'us; graph', round(d['cuda_graph']['value']/1e9,1), 'rollout', round(d['fused_rollout']['value']/1e9,1))"
  done
done
python -m gym_cellular_b200.build --force > /dev/null 2>&1
