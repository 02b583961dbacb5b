set -x
timeout -s KILL 600 python benchmarks/window_check.py 1048576 2>&1 | tail -8
for rep in 1 2; do
for o in window_levels=0,tape_global=0 window_levels=0,tape_global=1 window_levels=2,tape_global=0 window_levels=2,tape_global=1 window_levels=3,tape_global=0 window_levels=3,tape_global=1 window_levels=4,tape_global=0 window_levels=4,tape_global=1 \
         window_levels=4,tape_elems=8,tape_global=1 window_levels=4,tape_elems=8,tape_global=0 window_levels=3,window_ring_extra=2,tape_global=1; do
  FMC_OPTIONS=$o timeout -s KILL 300 python benchmarks/lmm_sim_only.py 1048576 2>&1 | tail -1
done
done
timeout -s KILL 300 python benchmarks/swaption_kernel_study.py 2>&1 | tail -8
