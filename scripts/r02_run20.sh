# new tests, then ncu captures of the stochastic 16 x 4 kernels (int8 and packed) and the 10 x 8 pair8 kernel
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x -k "packed_state or final_obs_and_side_effect_rows or rejected_tables" > gpurun_out/r02_tests20.log 2>&1
tail -3 gpurun_out/r02_tests20.log
python scripts/shape_sweep.py --only 6 > gpurun_out/r02_sweep20.log 2>&1
python scripts/shape_sweep.py --only 6 --packed >> gpurun_out/r02_sweep20.log 2>&1
python scripts/shape_sweep.py --only 8 >> gpurun_out/r02_sweep20.log 2>&1
cat gpurun_out/r02_sweep20.log
ncu --set full --clock-control none --import-source on -k regex:cell_pair_kernel -s 10 -c 1 -f -o gpurun_out/r02_prof_16x4_noise python scripts/shape_sweep.py --only 6 > gpurun_out/ncu_20a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cell_packed_kernel -s 10 -c 1 -f -o gpurun_out/r02_prof_16x4_noise_packed python scripts/shape_sweep.py --only 6 --packed > gpurun_out/ncu_20b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cell_pair8_kernel -s 10 -c 1 -f -o gpurun_out/r02_prof_10x8 python scripts/shape_sweep.py --only 8 > gpurun_out/ncu_20c.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
