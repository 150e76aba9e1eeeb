# round 2, GPU call 33: gc_step_many in one launch (small shards): parity tests, then configs 2 and 3 with it on / off
mkdir -p gpurun_out
python -m pytest tests/test_gpu_many.py -q --no-header -rf --timeout 900 -x > gpurun_out/r02_tests33.log 2>&1; tail -5 gpurun_out/r02_tests33.log
python -m pytest tests -m gpu -q --no-header -rf --timeout 900 -x -k "step_many or graph or bound" >> gpurun_out/r02_tests33.log 2>&1; tail -3 gpurun_out/r02_tests33.log
B="python bench.py --steps 2000 --warmup 20 --no-extra --no-cpu-baseline --no-side"
for w in cfg2 cfg3; do
  $B --workload $w > gpurun_out/r02_many_$w.json 2> gpurun_out/r02_many_$w.err; tail -c 600 gpurun_out/r02_many_$w.json; echo
  GC_B200_STEP_MANY_FUSED=0 $B --workload $w > gpurun_out/r02_many_off_$w.json 2> gpurun_out/r02_many_off_$w.err; tail -c 600 gpurun_out/r02_many_off_$w.json; echo
done
