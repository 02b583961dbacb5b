set -x
for p in 10000 100000 200000 400000 1048576 2097152; do for o in window_levels=3 window_levels=3,window_elems=8,window_cta_warps=8 window_levels=3,window_elems=8 window_levels=0; do FMC_OPTIONS=$o timeout -s KILL 300 python benchmarks/lmm_sim_only.py $p 2>&1 | tail -1; done; done
