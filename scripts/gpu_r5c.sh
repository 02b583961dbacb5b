# Brownian grid: one wave (occupancy query) against five and three blocks per SM; bit-exactness tests
for o in brownian_blocks_per_sm=5 brownian_blocks_per_sm=0 brownian_blocks_per_sm=3 brownian_blocks_per_sm=8; do echo "== $o"; for p in 1048576 100000 10000; do FMC_OPTIONS=$o timeout -s KILL 120 python benchmarks/brownian_rate.py $p 2>&1 | tail -1; done; done
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "brownian or mt19937 or slices" 2>&1 | tail -2
