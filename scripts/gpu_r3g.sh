timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_workloads.py -m gpu -x -q -k "regression or fuzzed_statistics or bermudan" 2>&1 | tail -3
timeout -s KILL 600 python benchmarks/raw_ops.py 2>&1 | grep -i -E "regress" | tail -12
