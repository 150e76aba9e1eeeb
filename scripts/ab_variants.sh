# A/B of compile-time kernel variants on ONE box (rebuilds on the GPU box with nvcc)
for v in "-DGC_GRID_SMALL_ITERS=0" "-DGC_GRID_SMALL_ITERS=100000"; do
  GC_NVCC_EXTRA="$v" python -m gym_cellular_b200.build --force > /dev/null 2>&1
  echo "== variant [$v]"
  python scripts/gw_size_sweep.py
done
python -m gym_cellular_b200.build --force > /dev/null 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
