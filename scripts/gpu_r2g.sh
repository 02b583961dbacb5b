# round 2: ncu of a typical LMM step kernel at 256 Ki paths (lone-warp regime), fused forms on
set -x
export FMC_OPTIONS=tape_elems=16
timeout -s KILL 300 python benchmarks/profile_lmm.py 262144 2 > gpurun_out/prof_plain.log 2>&1 && \
timeout -s KILL 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:tape_kernel -s 226 -c 3 -o gpurun_out/prof_sim256k_r2g -f python benchmarks/profile_lmm.py 262144 2 > gpurun_out/ncu_sim256k.log 2>&1
echo "ncu rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:tape_kernel -s 330 -c 3 -o gpurun_out/prof_swp256k_r2g -f python benchmarks/profile_lmm.py 262144 2 > gpurun_out/ncu_swp256k.log 2>&1
echo "ncu rc=$?"
