# round 2, first GPU call: full gpu test suite, packed-layout numbers, register-budget A/B, one ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r02_smi.txt
( time python -m pytest tests -m gpu -q --no-header -rf -x --timeout 900 ) > gpurun_out/r02_tests1.log 2>&1
tail -5 gpurun_out/r02_tests1.log
P="python scripts/packed_bench.py"
$P > gpurun_out/r02_packed_cfg4.json 2> gpurun_out/r02_packed_cfg4.err; cat gpurun_out/r02_packed_cfg4.json
$P --stochastic --no-host > gpurun_out/r02_packed_cfg4_stoch.json 2>&1; cat gpurun_out/r02_packed_cfg4_stoch.json
$P --envs 8388608 --cells 3 --levels 3 --no-host > gpurun_out/r02_packed_c3.json 2>&1; cat gpurun_out/r02_packed_c3.json
$P --envs 8388608 --cells 3 --levels 3 --no-host --stochastic > gpurun_out/r02_packed_c3s.json 2>&1; cat gpurun_out/r02_packed_c3s.json
$P --envs 65536 --cells 3 --levels 3 --no-host --steps 3000 > gpurun_out/r02_packed_cfg2.json 2>&1; cat gpurun_out/r02_packed_cfg2.json
ncu --set full --clock-control none --import-source on -k regex:cell_packed_kernel -s 12 -c 2 -f -o gpurun_out/r02_prof_packed $P --no-host --no-int8 --steps 20 > gpurun_out/ncu_packed.log 2>&1
for v in 3 5 6; do
  GC_NVCC_EXTRA="-DGC_PACKED_MINB=$v" python -m gym_cellular_b200.build --force > /dev/null 2>&1
  echo "MINB=$v"; $P --no-host --no-int8 2>&1 | tail -1
done
for v in 19 23; do
  GC_NVCC_EXTRA="-DGC_PACKED_BIG_ENVS=(1<<$v)" python -m gym_cellular_b200.build --force > /dev/null 2>&1
  echo "BIG_ENVS=1<<$v (cfg4 2^24: plain table when 1<<$v > 2^24)"; $P --no-host --no-int8 --envs 4194304 2>&1 | tail -1
done
