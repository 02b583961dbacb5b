for opts in "alternate_order=0" "alternate_order=1"; do
  echo "== $opts"
  FMC_OPTIONS=$opts timeout -s KILL 300 python benchmarks/lmm_phases.py 1048576 2>&1 | grep -E "kernels:|full step" | tail -4
done
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
