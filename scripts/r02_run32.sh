# round 2, GPU call 32: side-effect rows from the 5..8-level kernel (WITH_SE); before / after on the same box
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --no-header -rf --timeout 900 -x -k "pair8 or other_shapes or final_obs or generic" > gpurun_out/r02_tests32.log 2>&1; tail -3 gpurun_out/r02_tests32.log
for v in prev default; do
  if [ $v = default ]; then unset GC_B200_LIB_DIR; else export GC_B200_LIB_DIR=$PWD/gpu_variants/$v; fi
  echo "== $v" >> gpurun_out/r02_sweep32.log
  for c in 8 9 13 14; do python scripts/shape_sweep.py --only $c >> gpurun_out/r02_sweep32.log 2>&1; done
done
cat gpurun_out/r02_sweep32.log
