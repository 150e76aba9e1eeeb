# round 2, 8-GPU call: weak scaling of configs 4 and 5 with the statistics all-reduce timed, e2e against the
# box's aggregate PCIe ceiling, shard check on hardware
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/r02_topo_8gpu.txt
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 200 --warmup 5 > gpurun_out/r02_bench_${n}gpu.json 2> gpurun_out/r02_bench_${n}gpu.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_${n}gpu.json').read().strip().splitlines()[0])
w=d['workloads']['cfg5']
print('N=$n cfg4', round(d['value']/1e9,1), 'G frac/gpu', round(d['roofline']['frac'],4), 'packed', round(d['packed']['value']/1e9,1), 'e2e', round(d['e2e']['value']/1e9,2), 'link frac', round(d['e2e'].get('frac_of_link_ceiling',0),3), d['e2e']['pcie_measured'].get('all_ranks_concurrent_sum'), 'shard', d['shard_check'], d['episode_stats']['consistent'])
print('     cfg5', round(w['value']/1e9,1), 'G frac/gpu', round(w['roofline_frac'],4), 'e2e', round(w['e2e']/1e9,2), w['episode_stats_consistent'])
PY
done
python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_1gpu_samebox.json 2> gpurun_out/r02_bench_1gpu_samebox.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_1gpu_samebox.json').read().strip().splitlines()[0])
w=d['workloads']['cfg5']
print('N=1 cfg4', round(d['value']/1e9,1), 'G frac', round(d['roofline']['frac'],4), 'packed', round(d['packed']['value']/1e9,1), 'e2e', round(d['e2e']['value']/1e9,2), d['e2e']['pcie_measured'])
print('     cfg5', round(w['value']/1e9,1), 'G frac', round(w['roofline_frac'],4), 'e2e', round(w['e2e']/1e9,2))
PY
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_8gpubox.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_reference_8gpubox.json
( time python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -rf --timeout 900 ) > gpurun_out/r02_tests_8gpu.log 2>&1; tail -4 gpurun_out/r02_tests_8gpu.log
