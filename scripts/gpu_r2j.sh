set -x
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
set +x
for opts in "tape_elems=16" "tape_elems=16,tape_upload_stream=0" "tape_elems=16,flush_threshold=8192" "tape_elems=8"; do
  echo "== $opts"
  FMC_OPTIONS=$opts timeout -s KILL 300 python benchmarks/lmm_phases.py 1048576 2>&1 | grep -E "kernels:|full step|simulate:|swaption phase" | tail -7
done
