set -x
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash scripts/gpu_r2p.sh 2>/dev/null | grep "T="
