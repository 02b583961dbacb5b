timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "exp_log or unary or fused_chain or fuzzed_programs" 2>&1 | tail -3
timeout -s KILL 600 python benchmarks/raw_ops.py 2>&1 | grep -i -E "B1 exp|B1 log|euler" | tail -6
timeout -s KILL 1500 python benchmarks/explog_gpu_exhaustive.py 0 2>&1 | tail -4
