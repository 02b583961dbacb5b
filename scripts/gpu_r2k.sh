set -x
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
FMC_TEST_OPTIONS=tape_elems=8 timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_workloads.py -m gpu -x -q 2>&1 | tail -3
FMC_TEST_OPTIONS=tape_elems=4 timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_workloads.py -m gpu -x -q 2>&1 | tail -3
set +x
for opts in "tape_elems=16" "tape_elems=8"; do
  echo "== $opts"
  FMC_OPTIONS=$opts timeout -s KILL 300 python benchmarks/lmm_phases.py 1048576 2>&1 | grep -E "kernels:|full step|simulate:|swaption phase" | tail -7
done
