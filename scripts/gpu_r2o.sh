set -x
timeout -s KILL 900 python -m pytest tests/test_gpu_aad.py -m gpu -x -q 2>&1 | tail -12
timeout -s KILL 600 python benchmarks/aad_footprint.py 1048576 40 2>&1 | tail -5
