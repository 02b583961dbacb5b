set -x
timeout -s KILL 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for o in window_levels=3 window_levels=3,window_elems=8,window_cta_warps=8 window_levels=0; do
echo "== $o"
for rep in 1 2; do
FMC_OPTIONS=$o timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],3),'kernel',round(d['roofline']['kernel_ms_per_step'],3),'frac',round(d['roofline']['frac'],4),'e2e',round(d['e2e']['ms_per_step'],2),'pp',round(d['price_products_step']['ms_per_step'],2), 'host', d.get('host_profile'))"
done
done
timeout -s KILL 300 python benchmarks/swaption_kernel_study.py 1048576 4 2>&1 | grep -E "^n="
