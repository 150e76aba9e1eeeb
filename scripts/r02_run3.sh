# round 2, GPU call 3: tests after the wide-noise change, chain probe, shape sweeps (register budgets), packed numbers
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --no-header -rf --timeout 900 ) > gpurun_out/r02_tests3.log 2>&1
tail -8 gpurun_out/r02_tests3.log
python scripts/lane_probe.py --kind gridworld --envs 1048576 > gpurun_out/r02c_lanes_gw.json 2>&1; cat gpurun_out/r02c_lanes_gw.json
python scripts/lane_probe.py --kind cellular --envs 65536 > gpurun_out/r02c_lanes_c3.json 2>&1; cat gpurun_out/r02c_lanes_c3.json
python scripts/lane_probe.py --kind gridworld --envs 4194304 > gpurun_out/r02c_lanes_gw4m.json 2>&1; cat gpurun_out/r02c_lanes_gw4m.json
python scripts/shape_sweep.py > gpurun_out/r02c_shapes.txt 2>&1; cat gpurun_out/r02c_shapes.txt
python scripts/shape_sweep.py --cells > gpurun_out/r02c_shapes_cells.txt 2>&1; cat gpurun_out/r02c_shapes_cells.txt
P="python scripts/packed_bench.py --no-host"
$P > gpurun_out/r02c_packed_cfg4.json 2>&1; cat gpurun_out/r02c_packed_cfg4.json
$P --stochastic > gpurun_out/r02c_packed_cfg4s.json 2>&1; cat gpurun_out/r02c_packed_cfg4s.json
$P --cells 8 --stochastic > gpurun_out/r02c_packed_c8s.json 2>&1; cat gpurun_out/r02c_packed_c8s.json
