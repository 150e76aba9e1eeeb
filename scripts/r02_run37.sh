# round 2, GPU call 37: many-step kernels with the actions of two steps in flight, against one (gpu_variants/pf1)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_many.py -q --no-header -rf --timeout 900 -x > gpurun_out/r02_tests37.log 2>&1; tail -3 gpurun_out/r02_tests37.log
B="python bench.py --steps 2000 --warmup 20 --no-extra --no-cpu-baseline --no-side"
for v in pf1 default pf1 default; do
  if [ $v = default ]; then unset GC_B200_LIB_DIR; else export GC_B200_LIB_DIR=$PWD/gpu_variants/$v; fi
  for w in cfg2 cfg3; do
    $B --workload $w > gpurun_out/r02_pf_${v}_$w.json 2> gpurun_out/r02_pf_${v}_$w.err
    python -c "
import json,sys
d=json.loads(open('gpurun_out/r02_pf_${v}_$w.json').read().strip().splitlines()[-1])
print('$v $w', round(d['value']/1e9,1), round(d['ms_per_step']*1e3,3), 'packed', d.get('packed') and round(d['packed']['value']/1e9,1))"
  done
done
