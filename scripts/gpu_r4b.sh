set -x
timeout -s KILL 600 python benchmarks/window_check.py 1048576 2>&1 | tail -8
for o in window_levels=0 window_levels=2 window_levels=3 window_levels=4 window_levels=2,window_ring_extra=2 window_levels=3,window_ring_extra=2 window_levels=4,window_ring_extra=2 window_levels=3,window_ring_extra=4 \
         window_levels=3,tape_elems=8 window_levels=4,tape_elems=8 window_levels=5,tape_elems=8 window_levels=6,tape_elems=8 window_levels=4,tape_elems=8,window_ring_extra=2 \
         window_levels=4,cta_warps=6 window_levels=3,cta_warps=6 window_levels=4,flush_threshold=8192 window_levels=6,tape_elems=8,flush_threshold=8192; do
  FMC_OPTIONS=$o timeout -s KILL 300 python benchmarks/lmm_sim_only.py 1048576 2>&1 | tail -1
done
