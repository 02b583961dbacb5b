# round 2: is the LMM step at 1 Mi paths bound by the latency of a lone warp or by throughput? kernel times against the path count
for p in 1048576 524288 262144 131072; do
  for e in 16 8 4; do
    echo "== paths $p elems $e"
    FMC_OPTIONS=tape_elems=$e timeout -s KILL 300 python benchmarks/lmm_phases.py $p 2>&1 | grep "kernels:"
  done
done
