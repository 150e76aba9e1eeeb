# round 2, GPU call 41: int8 many-step kernel in 128-thread blocks for shards that do not fill the SMs, against 256 (gpu_variants/t256)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_many.py -q --no-header -rf --timeout 900 -x > gpurun_out/r02_tests41.log 2>&1; tail -2 gpurun_out/r02_tests41.log
B="python bench.py --steps 2000 --warmup 20 --no-extra --no-cpu-baseline --no-side"
for v in t256 default t256 default; do
  if [ $v = default ]; then unset GC_B200_LIB_DIR; else export GC_B200_LIB_DIR=$PWD/gpu_variants/$v; fi
  $B --workload cfg2 > gpurun_out/r02_t_${v}.json 2> gpurun_out/r02_t_${v}.err
  python -c "
import json
d=json.loads(open('gpurun_out/r02_t_${v}.json').read().strip().splitlines()[-1])
print('$v cfg2', round(d['value']/1e9,1), round(d['ms_per_step']*1e3,3), 'packed', round(d['packed']['value']/1e9,1))"
done
