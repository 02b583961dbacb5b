timeout -s KILL 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
for i in 1 2; do timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],2),'kernel_ms',round(d['roofline']['kernel_ms_per_step'],2),'e2e_ms',round(d['e2e']['ms_per_step'],1), 'frac', round(d['roofline']['frac'],3))"; done
timeout -s KILL 120 python benchmarks/lmm_sim_only.py 1048576 2>&1 | tail -1
