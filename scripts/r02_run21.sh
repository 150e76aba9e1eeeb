# wide-noise compares rewritten for the FMA pipe: full suite, then timings
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/r02_tests21.log 2>&1
tail -3 gpurun_out/r02_tests21.log
python scripts/shape_sweep.py --only 6 > gpurun_out/r02_sweep21.log 2>&1
python scripts/shape_sweep.py --only 6 --packed >> gpurun_out/r02_sweep21.log 2>&1
python scripts/shape_sweep.py --only 4 >> gpurun_out/r02_sweep21.log 2>&1
python scripts/shape_sweep.py --only 4 --packed >> gpurun_out/r02_sweep21.log 2>&1
python scripts/shape_sweep.py --only 9 >> gpurun_out/r02_sweep21.log 2>&1
cat gpurun_out/r02_sweep21.log
