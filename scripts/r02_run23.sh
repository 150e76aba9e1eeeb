# control: the committed wide-noise compare (gpu_variants/old) against the min-based tie detection, same box
mkdir -p gpurun_out
for v in old k0d0g0 default old; do
  if [ $v = default ]; then unset GC_B200_LIB_DIR; else export GC_B200_LIB_DIR=$PWD/gpu_variants/$v; fi
  echo "== $v" >> gpurun_out/r02_sweep23.log
  python scripts/shape_sweep.py --only 6 >> gpurun_out/r02_sweep23.log 2>&1
  python scripts/shape_sweep.py --only 6 --packed >> gpurun_out/r02_sweep23.log 2>&1
  python scripts/shape_sweep.py --only 4 >> gpurun_out/r02_sweep23.log 2>&1
  python scripts/shape_sweep.py --only 5 >> gpurun_out/r02_sweep23.log 2>&1
done
cat gpurun_out/r02_sweep23.log
