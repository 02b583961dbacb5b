timeout -s KILL 600 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout -s KILL 800 python scripts/berm_check.py 2>&1 | tail -5
timeout -s KILL 600 python benchmarks/configs.py 2>&1 | tail -5
