# round 2, GPU call 31: stochastic 5..8-level kernel: the new test, three or four blocks per SM, one ncu capture
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --no-header -rf --timeout 900 -x -k "pair8" > gpurun_out/r02_tests31.log 2>&1; tail -3 gpurun_out/r02_tests31.log
for v in default minb4 default minb4; do
  if [ $v = default ]; then unset GC_B200_LIB_DIR; else export GC_B200_LIB_DIR=$PWD/gpu_variants/$v; fi
  echo "== $v" >> gpurun_out/r02_sweep31.log
  python scripts/shape_sweep.py --only 9 >> gpurun_out/r02_sweep31.log 2>&1
done
unset GC_B200_LIB_DIR
cat gpurun_out/r02_sweep31.log
ncu --set full --clock-control none --import-source on -k regex:cell_pair8_kernel -s 10 -c 1 -f -o gpurun_out/r02_prof_10x8_noise python scripts/shape_sweep.py --only 9 > gpurun_out/ncu_31.log 2>&1
ls -la gpurun_out/r02_prof_10x8_noise.ncu-rep
