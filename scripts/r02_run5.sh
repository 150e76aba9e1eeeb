# round 2, GPU call 5: tests, L1-prefetch A/B of the grid-world kernel, sanitizer, profiles, bench lines
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --no-header -rf --timeout 900 ) > gpurun_out/r02_tests5.log 2>&1
tail -6 gpurun_out/r02_tests5.log
for i in 1 2; do python scripts/lane_probe.py --kind gridworld --envs 1048576 2>&1 | cut -c1-110; done
python bench.py --workload cfg3 --no-extra --no-cpu-baseline --steps 3000 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('cfg3 L1 prefetch', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step']*1e3,2), 'us; graph', round(d['cuda_graph']['value']/1e9,1))"
GC_NVCC_EXTRA="-DGC_GRID_NO_L1_PREFETCH" python -m gym_cellular_b200.build --force > /dev/null 2>&1
for i in 1 2; do python scripts/lane_probe.py --kind gridworld --envs 1048576 2>&1 | cut -c1-110; done
python bench.py --workload cfg3 --no-extra --no-cpu-baseline --steps 3000 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('cfg3 no prefetch', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step']*1e3,2), 'us; graph', round(d['cuda_graph']['value']/1e9,1))"
python -m gym_cellular_b200.build --force > /dev/null 2>&1
bash scripts/sanitize_smoke.sh
bash scripts/capture_profiles_r02.sh > gpurun_out/capture.log 2>&1; tail -3 gpurun_out/capture.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-300 gpurun_out/r02_bench_reference.json
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 1500 gpurun_out/r02_bench_default.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_driverlike.json 2> gpurun_out/r02_bench_driverlike.err; tail -c 600 gpurun_out/r02_bench_driverlike.json
