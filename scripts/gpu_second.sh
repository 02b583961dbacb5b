set -x
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python benchmarks/raw_ops.py --sizes 1048576,67108864 --out gpurun_out/raw_ops_r1c.json > gpurun_out/raw_ops_r1c.log 2>&1; echo "raw rc=$?"
timeout -s KILL 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/bench_r1c.json
FMC_LOG_TAPES=1 timeout -s KILL 300 python benchmarks/profile_cases.py > gpurun_out/prof_plain.log 2> gpurun_out/prof_tapes.log && \
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1c.csv python benchmarks/profile_cases.py > gpurun_out/ncu_launch.log 2>&1 && \
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:tape_kernel -s 6 -c 12 -o gpurun_out/prof_tape_r1c -f python benchmarks/profile_cases.py > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full.log
