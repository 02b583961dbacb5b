# round 2: CTA size / ring depth knobs at 1 Mi paths (balance of a one-wave grid over the 148 SMs)
for opts in "tape_elems=16,cta_warps=4" "tape_elems=16,cta_warps=2" "tape_elems=16,cta_warps=1" "tape_elems=8,cta_warps=4" "tape_elems=8,cta_warps=2" "tape_elems=8,cta_warps=8" "tape_elems=16,cta_warps=2,ring_max=4" "tape_elems=16,cta_warps=2,max_regs=4" "tape_elems=8,cta_warps=2,ring_max=4"; do
  echo "== $opts"
  FMC_OPTIONS=$opts timeout -s KILL 300 python benchmarks/lmm_phases.py 1048576 2>&1 | grep -E "kernels:" | tail -3
done
