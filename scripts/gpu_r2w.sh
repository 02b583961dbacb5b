timeout -s KILL 600 python benchmarks/ab_option.py alternate_order 0 1 1048576 6 2>&1 | tail -2
timeout -s KILL 600 python benchmarks/ab_option.py flush_threshold 4096 6144 1048576 6 2>&1 | tail -2
