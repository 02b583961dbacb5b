timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout -s KILL 300 python benchmarks/raw_ops.py --sizes 67108864 --cases ref --out gpurun_out/raw_ref.json 2>&1 | grep REF
