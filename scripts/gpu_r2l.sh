set -x
timeout -s KILL 1200 python -m pytest tests/test_gpu_workloads.py -m gpu -x -q --durations=5 2>&1 | tail -14
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2l.json 2> gpurun_out/bench_r2l.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2l.err
