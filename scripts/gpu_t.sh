timeout -s KILL 600 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout -s KILL 300 python benchmarks/raw_ops.py --sizes 1000000,100000000 --cases b4 --out gpurun_out/raw_b4.json 2>&1 | grep B4
