# A/B of the wide-noise compare variants (gpu_variants/<name>: K_TOP branch, IMAD d, 64-bit gather)
mkdir -p gpurun_out
for v in default k0d0g0 k1d0g0 k0d1g0 k0d0g1 k0d1g1; do
  if [ $v = default ]; then unset GC_B200_LIB_DIR; else export GC_B200_LIB_DIR=$PWD/gpu_variants/$v; fi
  echo "== $v" >> gpurun_out/r02_sweep22.log
  python scripts/shape_sweep.py --only 6 >> gpurun_out/r02_sweep22.log 2>&1
  python scripts/shape_sweep.py --only 6 --packed >> gpurun_out/r02_sweep22.log 2>&1
  python scripts/shape_sweep.py --only 4 >> gpurun_out/r02_sweep22.log 2>&1
done
cat gpurun_out/r02_sweep22.log
