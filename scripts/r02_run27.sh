# A/B: three groups of rows in flight before the draws (int8 wide stochastic), packed Philox kernel at four blocks
mkdir -p gpurun_out
for v in old m3 m3_p3 m3_p3_pk4 k1_p3_pk4; do
  export GC_B200_LIB_DIR=$PWD/gpu_variants/$v
  echo "== $v" >> gpurun_out/r02_sweep27.log
  python scripts/shape_sweep.py --only 6 >> gpurun_out/r02_sweep27.log 2>&1
  python scripts/shape_sweep.py --only 6 --packed >> gpurun_out/r02_sweep27.log 2>&1
  python scripts/shape_sweep.py --only 4 >> gpurun_out/r02_sweep27.log 2>&1
done
cat gpurun_out/r02_sweep27.log
