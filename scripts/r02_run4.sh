# round 2, GPU call 4: tests after the 16-bit-half noise draws, cell-count sweep, grid-world partition A/B
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --no-header -rf --timeout 900 ) > gpurun_out/r02_tests4.log 2>&1
tail -8 gpurun_out/r02_tests4.log
python scripts/shape_sweep.py --cells > gpurun_out/r02d_shapes_cells.txt 2>&1; cat gpurun_out/r02d_shapes_cells.txt
P="python scripts/packed_bench.py --no-host"
$P --stochastic > gpurun_out/r02d_packed_cfg4s.json 2>&1; cat gpurun_out/r02d_packed_cfg4s.json
$P --cells 8 --stochastic > gpurun_out/r02d_packed_c8s.json 2>&1; cat gpurun_out/r02d_packed_c8s.json
for i in 1 2; do python scripts/lane_probe.py --kind gridworld --envs 1048576 2>&1 | cut -c1-120; done
python bench.py --workload cfg3 --no-extra --no-cpu-baseline --steps 3000 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('cfg3 contiguous', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step']*1e3,2), 'us; graph', round(d['cuda_graph']['value']/1e9,1))"
python bench.py --workload cfg5 --no-extra --no-cpu-baseline --steps 2000 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('cfg5 contiguous', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step']*1e3,2), 'us frac', round(d['roofline']['frac'],3))"
GC_NVCC_EXTRA="-DGC_GRID_STRIDE_PARTITION" python -m gym_cellular_b200.build --force > /dev/null 2>&1
for i in 1 2; do python scripts/lane_probe.py --kind gridworld --envs 1048576 2>&1 | cut -c1-120; done
python bench.py --workload cfg3 --no-extra --no-cpu-baseline --steps 3000 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('cfg3 grid-stride', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step']*1e3,2), 'us; graph', round(d['cuda_graph']['value']/1e9,1))"
python bench.py --workload cfg5 --no-extra --no-cpu-baseline --steps 2000 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('cfg5 grid-stride', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step']*1e3,2), 'us frac', round(d['roofline']['frac'],3))"
