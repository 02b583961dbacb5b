for opts in "flush_threshold=4096" "flush_threshold=6144" "flush_threshold=8192" "flush_threshold=16384" "flush_threshold=3200"; do
  echo "== $opts"
  FMC_OPTIONS=$opts timeout -s KILL 300 python benchmarks/lmm_phases.py 1048576 2>&1 | grep -E "kernels:|full step|simulate:" | tail -5
done
