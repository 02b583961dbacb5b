# round 2: fused one-dispatch forms (RATIOACC, AXPYST, ADDAFFDISC + reload): parity, then kernel times against the path count
set -x
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for e in 16 8; do
  FMC_TEST_OPTIONS=tape_elems=$e timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_workloads.py -m gpu -x -q 2>&1 | tail -3
done
set +x
for p in 1048576 262144; do
  for e in 16 8; do
    for f in 1 0; do
      echo "== paths $p elems $e fuse_ops2 $f"
      FMC_OPTIONS=tape_elems=$e,fuse_ops2=$f timeout -s KILL 300 python benchmarks/lmm_phases.py $p 2>&1 | grep -E "kernels:|full step|swaption phase" | tail -5
    done
  done
done
