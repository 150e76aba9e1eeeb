# A/B of the tie detection (AND-accumulate / two-input lane minimum) and of the compare forms, against gpu_variants/old
mkdir -p gpurun_out
for v in old and_k0d0g0 min2_k0d0g0 and_k1d0g0 and_k1d0g1 and_k1d1g1; do
  export GC_B200_LIB_DIR=$PWD/gpu_variants/$v
  echo "== $v" >> gpurun_out/r02_sweep24.log
  python scripts/shape_sweep.py --only 6 >> gpurun_out/r02_sweep24.log 2>&1
  python scripts/shape_sweep.py --only 6 --packed >> gpurun_out/r02_sweep24.log 2>&1
  python scripts/shape_sweep.py --only 4 >> gpurun_out/r02_sweep24.log 2>&1
done
cat gpurun_out/r02_sweep24.log
