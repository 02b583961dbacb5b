timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 600 python benchmarks/raw_ops.py 2>&1 | grep -E "n=  67108864" | grep -i -E "B1 add|B1 mult|payoff|chain|getAverage|exp|log" | head -12
for i in 1 2; do timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],3),'kernel',round(d['roofline']['kernel_ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],2),'pp',round(d['price_products_step']['ms_per_step'],2))"; done
