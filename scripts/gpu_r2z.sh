set -x
timeout -s KILL 300 python benchmarks/profile_cases.py 4096 > gpurun_out/prof_cases_plain.log 2>&1 && \
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:tape_kernel -c 6 -o gpurun_out/prof_micro_r3m -f python benchmarks/profile_cases.py 4096 > gpurun_out/ncu_micro_r3m.log 2>&1
echo "ncu rc=$?"
