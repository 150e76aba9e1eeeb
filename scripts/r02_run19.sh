# round 2, GPU call 19: the gpu suite with the run-time switches flipped (PDL off, graph replay off, TMA variant on)
mkdir -p gpurun_out
for env in "GC_B200_PDL=0" "GC_B200_STEP_MANY_GRAPH=0" "GC_B200_TMA=1"; do
  ( export $env; python -m pytest tests -m gpu -q --no-header -rf --timeout 900 -x ) > gpurun_out/r02_tests19_$env.log 2>&1
  echo "$env: $(grep -E 'passed|failed' gpurun_out/r02_tests19_$env.log)"
done
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_driverlike.json 2> gpurun_out/r02_bench_driverlike.err; tail -c 150 gpurun_out/r02_bench_driverlike.json
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 150 gpurun_out/r02_bench_default.json
