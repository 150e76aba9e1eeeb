# A/B: L2 prefetch of the next word's rows in the wide stochastic int8 kernel
mkdir -p gpurun_out
for v in old m3 m3_pf m3_g1_pf m3_pf_regpf; do
  export GC_B200_LIB_DIR=$PWD/gpu_variants/$v
  echo "== $v" >> gpurun_out/r02_sweep26.log
  python scripts/shape_sweep.py --only 6 >> gpurun_out/r02_sweep26.log 2>&1
  python scripts/shape_sweep.py --only 4 >> gpurun_out/r02_sweep26.log 2>&1
done
cat gpurun_out/r02_sweep26.log
