# round 2, GPU call 39: int8 many-step kernel with the pair table replicated (32 KB), against the plain table (gpu_variants/norep)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_many.py -q --no-header -rf --timeout 900 -x > gpurun_out/r02_tests39.log 2>&1; tail -2 gpurun_out/r02_tests39.log
B="python bench.py --steps 2000 --warmup 20 --no-extra --no-cpu-baseline --no-side"
for v in norep default norep default; do
  if [ $v = default ]; then unset GC_B200_LIB_DIR; else export GC_B200_LIB_DIR=$PWD/gpu_variants/$v; fi
  $B --workload cfg2 > gpurun_out/r02_rep_${v}.json 2> gpurun_out/r02_rep_${v}.err
  python -c "
import json
d=json.loads(open('gpurun_out/r02_rep_${v}.json').read().strip().splitlines()[-1])
print('$v cfg2', round(d['value']/1e9,1), round(d['ms_per_step']*1e3,3), 'packed', round(d['packed']['value']/1e9,1))"
done
unset GC_B200_LIB_DIR
python scripts/shape_sweep.py --only 1 > /dev/null 2>&1
