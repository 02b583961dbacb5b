timeout -s KILL 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "regression_normal or batched or ragged or reductions or exp_log or fused_chain or compound" 2>&1 | tail -15
echo "sanitizer rc=$?"
