# programmatic dependent launch on / off, same build
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for v in 1 0 1 0; do
  echo "== GC_B200_PDL=$v"
  GC_B200_PDL=$v bash scripts/quick_bench.sh
done
