timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-calibration --no-extras 2>gpurun_out/bench_r3b.err | tee gpurun_out/bench_r3b.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],3),'kernel',round(d['roofline']['kernel_ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],2), d['host_profile']); print(d['price_products_step'])"
tail -3 gpurun_out/bench_r3b.err
