# round 2, GPU call 10: captures for profiles/ (<= 64 MiB), final bench lines, host-path chunk sweep
mkdir -p gpurun_out
bash scripts/capture_profiles_r02.sh > gpurun_out/capture.log 2>&1; tail -2 gpurun_out/capture.log; du -sh gpurun_out
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-160 gpurun_out/r02_bench_reference.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_driverlike.json 2> gpurun_out/r02_bench_driverlike.err; tail -c 300 gpurun_out/r02_bench_driverlike.json
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 300 gpurun_out/r02_bench_default.json
for c in 524288 1048576 2097152 4194304; do python scripts/packed_bench.py --no-int8 --steps 50 --chunk $c 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('host packed chunk', d['host_packed']['chunk'], round(d['host_packed']['ms_per_step'],3), 'ms', round(d['host_packed']['env_steps_per_s']/1e9,2), 'G')"; done
du -sh gpurun_out
