timeout -s KILL 1700 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 600 python benchmarks/raw_ops.py 2>&1 | grep -E "n= *100000000|n=  67108864" | grep -i -E "payoff|div|discount|chain" | head -12
