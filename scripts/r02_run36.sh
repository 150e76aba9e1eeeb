# round 2, GPU call 36: packed many-step kernel; packed_word shared by the step and many-step kernels (bench lines for the packed device path)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_many.py tests/test_gpu_packed.py -q --no-header -rf --timeout 900 -x > gpurun_out/r02_tests36.log 2>&1; tail -4 gpurun_out/r02_tests36.log
python bench.py --no-cpu-baseline > gpurun_out/r02_bench36.json 2> gpurun_out/r02_bench36.err
python scripts/shape_sweep.py --only 6 --packed > gpurun_out/r02_sweep36.log 2>&1
python scripts/shape_sweep.py --only 5 --packed >> gpurun_out/r02_sweep36.log 2>&1
python scripts/shape_sweep.py --only 1 --packed >> gpurun_out/r02_sweep36.log 2>&1
cat gpurun_out/r02_sweep36.log
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench36.json').read().strip().splitlines()[-1])
print('cfg4', round(d['value']/1e9,2), 'packed', round(d['packed']['value']/1e9,1), d['packed']['roofline']['frac'], 'e2e', round(d['e2e']['value']/1e9,2))
for k,v in d['workloads'].items():
    print(k, round(v['value']/1e9,1), 'packed', v['packed'] and round(v['packed']['value']/1e9,1), v['episode_stats_consistent'])
PY
