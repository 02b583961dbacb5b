for o in "" "fuse_ops=0" "ring_max=2" "ring_max=3" "ring_max=4,target_ctas=3" "ring_max=6,target_ctas=3" "ring_max=8,target_ctas=2" "ring_max=12,target_ctas=2" "pipeline=0" "cta_warps=2" "horizon=8"; do
  FMC_OPTIONS="$o" timeout -s KILL 120 python benchmarks/lmm_sim_only.py 2>&1 | tail -1
done
for p in 262144 4194304; do timeout -s KILL 120 python benchmarks/lmm_sim_only.py $p 2>&1 | tail -1; done
