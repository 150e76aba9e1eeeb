# round 2, 2-GPU call at HEAD: the driver-like line at N = 2 and the reference arm under torchrun
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_2gpu_driverlike.json 2> gpurun_out/r02_bench_2gpu_driverlike.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_2gpu_reference.json 2> gpurun_out/r02_bench_2gpu_reference.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_2gpu_driverlike.json').read().strip().splitlines()[-1]); w=d['workloads']['cfg5']
print('cfg4', round(d['value']/1e9,1), round(d['roofline']['frac'],4), 'packed', round(d['packed']['value']/1e9,1), 'cfg5', round(w['value']/1e9,1), round(w['roofline_frac'],4), 'e2e', round(d['e2e']['value']/1e9,2), d.get('shard_check'), d['episode_stats']['consistent'], w['episode_stats_consistent'])
r=json.loads(open('gpurun_out/r02_bench_2gpu_reference.json').read().strip().splitlines()[-1])
print('reference arm', r.get('impl'), round(r['value']/1e6,1), 'M env-steps/s', r['cpu_baseline']['cores'], 'cores')
PY
