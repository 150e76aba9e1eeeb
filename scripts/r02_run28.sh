# A/B: all row words of an iteration staged with cp.async into per-thread shared-memory slots (wide stochastic int8 kernel)
mkdir -p gpurun_out
for v in old m3 m3_as m3_as4; do
  export GC_B200_LIB_DIR=$PWD/gpu_variants/$v
  echo "== $v" >> gpurun_out/r02_sweep28.log
  python scripts/shape_sweep.py --only 6 >> gpurun_out/r02_sweep28.log 2>&1
  python scripts/shape_sweep.py --only 4 >> gpurun_out/r02_sweep28.log 2>&1
done
cat gpurun_out/r02_sweep28.log
export GC_B200_LIB_DIR=$PWD/gpu_variants/m3_as4
python -m pytest tests -q -m gpu -x > gpurun_out/r02_tests28.log 2>&1
tail -3 gpurun_out/r02_tests28.log
