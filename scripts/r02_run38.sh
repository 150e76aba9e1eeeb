# round 2, GPU call 38: the gpu suite with the run-time switches flipped (one-launch gc_step_many off, PDL off, graph replay off), then as shipped
mkdir -p gpurun_out
for env in "GC_B200_STEP_MANY_FUSED=0" "GC_B200_PDL=0" "GC_B200_STEP_MANY_GRAPH=0"; do
  ( export $env; python -m pytest tests -m gpu -q --no-header -rf --timeout 900 -x ) > gpurun_out/r02_tests38_$env.log 2>&1
  echo "$env: $(grep -E 'passed|failed' gpurun_out/r02_tests38_$env.log)"
done
python -m pytest tests -m gpu -q --no-header -rf --timeout 900 -x > gpurun_out/r02_tests38.log 2>&1; tail -2 gpurun_out/r02_tests38.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
