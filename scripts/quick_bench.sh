# every workload once: plain stepping, CUDA graph and fused rollout where they apply
for w in cfg2 cfg3; do
  python bench.py --workload $w --steps 3000 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$w', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step']*1e3,2), 'us; graph', round(d['cuda_graph']['value']/1e9,1), round(d['cuda_graph']['ms_per_step']*1e3,2), 'us; rollout', round(d['fused_rollout']['value']/1e9,1))"
done
for w in cfg4 cfg5; do
  python bench.py --workload $w --steps 2000 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$w', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step']*1e3,1), 'us frac', round(d['roofline']['frac'],3))"
done
