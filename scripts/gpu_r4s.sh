for o in window_cta_warps=8 window_cta_warps=12 window_cta_warps=16 window_cta_warps=20 window_cta_warps=24 window_levels=4,window_cta_warps=16 window_levels=4,window_cta_warps=20 window_elems=16,window_cta_warps=12 window_elems=16,window_cta_warps=10; do
  FMC_OPTIONS=$o timeout -s KILL 300 python benchmarks/lmm_sim_only.py 1048576 2>&1 | tail -1
done
FMC_LOG_TAPES=1 FMC_OPTIONS=window_cta_warps=20 timeout -s KILL 300 python benchmarks/lmm_sim_only.py 1048576 2>&1 | grep "fmc tape" | sed -n 3,5p
