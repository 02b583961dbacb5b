timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 120 python benchmarks/dispatch_latency.py 1 2>&1 | tail -6
for p in 262144 1048576 4194304; do timeout -s KILL 120 python benchmarks/lmm_sim_only.py $p 2>&1 | tail -1; done
timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],2),'kernel_ms',round(d['roofline']['kernel_ms_per_step'],2),'e2e_ms',round(d['e2e']['ms_per_step'],1), 'frac', round(d['roofline']['frac'],3))"
timeout -s KILL 300 python benchmarks/raw_ops.py --sizes 67108864 --cases b1,b2 --out gpurun_out/raw_ops_r1f.json 2>&1 | grep -E "B1|B2" | cut -c1-110
