timeout -s KILL 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout -s KILL 600 python benchmarks/configs.py 2>&1 | tail -5
timeout -s KILL 300 python benchmarks/raw_ops.py --sizes 1000000 --cases b5 --out gpurun_out/raw_b5.json 2>&1 | grep B5
