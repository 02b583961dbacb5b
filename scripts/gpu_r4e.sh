set -x
for o in window_levels=3 window_levels=3,window_elems=8,window_cta_warps=8 window_levels=3,window_elems=8 window_levels=4,window_elems=8,window_cta_warps=8 window_levels=3,window_cta_warps=8; do
echo "== $o"
for rep in 1 2; do
FMC_OPTIONS=$o timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],3),'kernel',round(d['roofline']['kernel_ms_per_step'],3),'frac',round(d['roofline']['frac'],4),'e2e',round(d['e2e']['ms_per_step'],2),'pp',round(d['price_products_step']['ms_per_step'],2), 'host', d.get('host_profile'))"
done
done
for m in 4 20; do for o in tape_elems=16 tape_elems=8 tape_elems=8,cta_warps=8 tape_elems=16,cta_warps=8 tape_elems=16,cta_warps=2; do echo "== m=$m $o"; FMC_OPTIONS=$o timeout -s KILL 300 python benchmarks/swaption_kernel_study.py 1048576 $m 2>&1 | grep -E "full|empty|sum of"; done; done
timeout -s KILL 300 python benchmarks/latency.py 2>&1 | tail -9
