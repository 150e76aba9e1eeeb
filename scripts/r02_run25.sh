# A/B after the rare tie path stopped reading the round keys through a pointer
mkdir -p gpurun_out
for v in old and_k0d0g0 min3_k0d0g0 and_k1d1g1 min3_k1d1g1 min3_k1d0g1 min3_k1d0g0 min3_k0d0g1; do
  export GC_B200_LIB_DIR=$PWD/gpu_variants/$v
  echo "== $v" >> gpurun_out/r02_sweep25.log
  python scripts/shape_sweep.py --only 6 >> gpurun_out/r02_sweep25.log 2>&1
  python scripts/shape_sweep.py --only 6 --packed >> gpurun_out/r02_sweep25.log 2>&1
  python scripts/shape_sweep.py --only 4 >> gpurun_out/r02_sweep25.log 2>&1
done
cat gpurun_out/r02_sweep25.log
