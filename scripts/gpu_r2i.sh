set -x
export FMC_OPTIONS=tape_elems=16
timeout -s KILL 300 python benchmarks/profile_lmm.py 1048576 2 > gpurun_out/prof_plain.log 2>&1 && \
timeout -s KILL 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:tape_kernel -s 227 -c 3 -o gpurun_out/prof_sim1m_r2i -f python benchmarks/profile_lmm.py 1048576 2 > gpurun_out/ncu_sim1m.log 2>&1
echo "ncu rc=$?"
