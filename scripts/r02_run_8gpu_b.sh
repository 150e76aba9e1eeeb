# 8-GPU: short and long runs of config 4 with the device-side start alignment
mkdir -p gpurun_out
for spec in "20 5" "20 5" "200 5"; do
  set -- $spec
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps $1 --warmup $2 --no-extra > gpurun_out/r02b_bench_8gpu_$1.json 2> gpurun_out/r02b_bench_8gpu_$1.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02b_bench_8gpu_$1.json').read().strip().splitlines()[0])
print('N=8 steps=$1 cfg4', round(d['value']/1e9,1), 'G us/step', round(d['ms_per_step']*1e3,1), 'frac/gpu', round(d['roofline']['frac'],4), 'packed', round(d['packed']['value']/1e9,1), round(d['packed']['ms_per_step']*1e3,1), 'e2e', round(d['e2e']['value']/1e9,2), 'shard', d['shard_check'])
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_8gpu_driverlike.json 2> gpurun_out/r02_bench_8gpu_driverlike.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_8gpu_driverlike.json').read().strip().splitlines()[0])
w=d['workloads']['cfg5']
print('N=8 driver-like cfg4', round(d['value']/1e9,1), 'G frac/gpu', round(d['roofline']['frac'],4), 'cfg5', round(w['value']/1e9,1), round(w['roofline_frac'],4), 'e2e', round(d['e2e']['value']/1e9,2))
PY
