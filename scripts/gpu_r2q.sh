set -x
timeout -s KILL 300 python benchmarks/profile_brownian.py > gpurun_out/prof_plain.log 2>&1 && \
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:brownian_kernel -s 1 -c 1 -o gpurun_out/prof_brownian_r2q -f python benchmarks/profile_brownian.py > gpurun_out/ncu_brownian.log 2>&1
echo "ncu rc=$?"
