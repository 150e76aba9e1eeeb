# A/B of compile-time kernel variants over every cell count (rebuilds on the GPU box with nvcc)
for v in "" "-DGC_PAIR_MINB=3"; do
  GC_NVCC_EXTRA="$v" python -m gym_cellular_b200.build --force > /dev/null 2>&1
  echo "== variant [$v]"
  python scripts/shape_sweep.py --cells | awk '{print $2, $6, $(NF-8), $(NF-7), $(NF-6), $(NF-5)}'
done
python -m gym_cellular_b200.build --force > /dev/null 2>&1
for w in cfg5 cfg4; do python bench.py --workload $w --steps 2000 --warmup 20 --no-extra --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(\"$w\", d[\"value\"], d[\"roofline\"][\"frac\"])"; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
