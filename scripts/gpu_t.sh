timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout -s KILL 300 python benchmarks/raw_ops.py --sizes 67108864 --cases b1,b2 --out gpurun_out/raw_ops_r1e.json 2>&1 | grep -E "div|discount|payoff"
