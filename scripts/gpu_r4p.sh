for p in 10000 1048576; do for o in window_levels=3; do echo "== $p $o"; for rep in 1 2; do FMC_OPTIONS=$o timeout -s KILL 300 python benchmarks/lmm_phases.py $p 2>&1 | grep -A3 "timing off" | tail -3; done; done; done
timeout -s KILL 900 python -m pytest tests/test_gpu_workloads.py tests/test_gpu_lifecycle.py -m gpu -x -q 2>&1 | tail -3
