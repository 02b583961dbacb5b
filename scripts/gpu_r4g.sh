set -x
for o in zero_copy_reduce=1 zero_copy_reduce=0 grid_limit=148 grid_limit=296 cta_warps=8 pipeline=0; do echo "== $o"; FMC_OPTIONS=$o timeout -s KILL 300 python benchmarks/swaption_kernel_study.py 1048576 4 2>&1 | grep -E "^n=" | grep -E "full|empty|leaf"; done
