# compute-sanitizer over __graft_entry__.smoke() (both layouts, grid world, host path), PDL on and off
mkdir -p gpurun_out
for pdl in 1 0; do
  for tool in memcheck racecheck; do
    GC_B200_PDL=$pdl timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" \
      > gpurun_out/r02_sanitizer_${tool}_pdl${pdl}.log 2>&1
    echo "sanitizer tool=$tool PDL=$pdl exit=$?"; tail -3 gpurun_out/r02_sanitizer_${tool}_pdl${pdl}.log
  done
done
