# round 2, GPU call 14: kernels per gc_step_many graph, configs 2 and 3
show='import json,sys; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d["value"]/1e9,2), "G/s", round(d["ms_per_step"]*1e3,3), "us")'
for k in 8 16 32 64; do
  for w in cfg2 cfg3; do
    GC_B200_STEP_MANY_GRAPH_STEPS=$k python bench.py --workload $w --steps 2048 --no-extra --no-cpu-baseline --no-side 2>/dev/null | python -c "$show" "$w graph of $k steps"
  done
done
