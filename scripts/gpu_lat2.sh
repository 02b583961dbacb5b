timeout -s KILL 120 python benchmarks/dispatch_latency.py 1 2>&1 | tail -4
timeout -s KILL 120 python benchmarks/dispatch_latency.py 4 2>&1 | tail -4
