for o in window_levels=3 window_levels=4,window_cta_warps=16 window_levels=3,window_cta_warps=16; do
echo "== $o"
for rep in 1 2; do FMC_OPTIONS=$o timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],3),'kernel',round(d['roofline']['kernel_ms_per_step'],3),'frac',round(d['roofline']['frac'],4),'e2e',round(d['e2e']['ms_per_step'],2),'pp',round(d['price_products_step']['ms_per_step'],2))"; done
for p in 10000 200000; do FMC_OPTIONS=$o timeout -s KILL 300 python benchmarks/lmm_phases.py $p 2>&1 | grep -A3 "timing off" | grep "full step"; done
done
