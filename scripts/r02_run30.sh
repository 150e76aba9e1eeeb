# round 2, GPU call 30: stochastic variant of the 5..8-level kernel (single-cell table with the fire bit)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --no-header -rf --timeout 900 -x -k "pair8 or other_shapes or final_obs" > gpurun_out/r02_tests30.log 2>&1; tail -5 gpurun_out/r02_tests30.log
python scripts/shape_sweep.py --only 9 > gpurun_out/r02_sweep30.log 2>&1
python scripts/shape_sweep.py --only 8 >> gpurun_out/r02_sweep30.log 2>&1
cat gpurun_out/r02_sweep30.log
