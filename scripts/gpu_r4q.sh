for m in 4 20; do timeout -s KILL 300 python benchmarks/kernel_timeline.py 1048576 $m 2>&1 | tail -12; done
timeout -s KILL 300 python benchmarks/kernel_timeline.py 10000 20 2>&1 | tail -12
