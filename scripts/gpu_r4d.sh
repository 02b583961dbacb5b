set -x
for o in window_levels=0 window_levels=3 window_levels=3,cta_warps=8 window_levels=2,cta_warps=8 window_levels=3,tape_elems=8,cta_warps=8 window_levels=4,tape_elems=8,cta_warps=8 window_levels=4,tape_elems=8 window_levels=5,tape_elems=8,cta_warps=8 \
   window_levels=3,window_ring_extra=2 window_levels=3,window_ring_extra=4 window_levels=3,flush_threshold=8192 window_levels=3,min_warps=3; do
  FMC_OPTIONS=$o timeout -s KILL 300 python benchmarks/lmm_sim_only.py 1048576 2>&1 | tail -1
done
for o in window_levels=0 window_levels=3 window_levels=2 window_levels=4,tape_elems=8; do
echo "== $o"
FMC_OPTIONS=$o timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],3),'kernel',round(d['roofline']['kernel_ms_per_step'],3),'frac',d['roofline']['frac'],'e2e',round(d['e2e']['ms_per_step'],2),'pp',round(d['price_products_step']['ms_per_step'],2), 'parity', d.get('parity'))"
done
for p in 10000 100000 400000; do for o in window_levels=0 window_levels=3; do FMC_OPTIONS=$o timeout -s KILL 300 python benchmarks/lmm_sim_only.py $p 2>&1 | tail -1; done; done
