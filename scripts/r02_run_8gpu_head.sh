# round 2, 8-GPU call at HEAD: driver-like and 200-step lines at N = 8
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_8gpu_driverlike.json 2> gpurun_out/r02_bench_8gpu_driverlike.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 200 --warmup 5 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
python - <<PY
import json
for f in ('r02_bench_8gpu_driverlike','r02_bench_8gpu'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); w=d['workloads']['cfg5']
    print(f, 'cfg4', round(d['value']/1e9,1), 'frac/gpu', round(d['roofline']['frac'],4), 'packed', round(d['packed']['value']/1e9,1), 'cfg5', round(w['value']/1e9,1), round(w['roofline_frac'],4), 'e2e', round(d['e2e']['value']/1e9,2), round(d['e2e'].get('frac_of_link_ceiling',0),3), d.get('shard_check'), d['episode_stats']['consistent'], w['episode_stats_consistent'])
PY
