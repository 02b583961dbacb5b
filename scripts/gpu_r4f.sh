set -x
for m in 4 20; do echo "== m=$m"; timeout -s KILL 300 python benchmarks/swaption_kernel_study.py 1048576 $m 2>&1 | grep -E "^n="; done
