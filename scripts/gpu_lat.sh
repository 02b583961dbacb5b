timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 120 python benchmarks/dispatch_latency.py 1 2>&1 | tail -7
timeout -s KILL 120 python benchmarks/dispatch_latency.py 4 2>&1 | tail -7
for p in 262144 1048576 4194304; do timeout -s KILL 120 python benchmarks/lmm_sim_only.py $p 2>&1 | tail -1; done
