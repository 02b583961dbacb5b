# A/B of window_reduce_min (512 vs 2048) on the working tree: bench step and the AAD footprint benchmark
for o in window_reduce_min=2048 window_reduce_min=512 window_reduce_min=2048; do
echo "== $o"
FMC_OPTIONS=$o timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],3),'kernel',round(d['roofline']['kernel_ms_per_step'],3),'frac',round(d['roofline']['frac'],4),'e2e',round(d['e2e']['ms_per_step'],2),'pp',round(d['price_products_step']['ms_per_step'],2),'launches',d.get('gpu_launches'))"
FMC_OPTIONS=$o timeout -s KILL 600 python benchmarks/aad_footprint.py 2>&1 | tail -4 | cut -c1-260
done
