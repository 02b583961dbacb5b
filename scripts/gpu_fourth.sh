timeout -s KILL 300 python benchmarks/latency.py 2>&1 | tail -12
timeout -s KILL 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],2),'kernel_ms',round(d['roofline']['kernel_ms_per_step'],2),'launches',d['gpu_launches_per_step'],'e2e_ms',round(d['e2e']['ms_per_step'],1), d['host_profile'], 'frac', round(d['roofline']['frac'],3))"
timeout -s KILL 120 python - <<'P'
import sys
sys.path[:0]=['.','finmath-lib-cuda-extensions_b200']
import finmath_cuda as fc, numpy as np
from finmath_cuda import _capi as capi
fc.ensure_init(0)
n=1<<26
for name,arr in (("ones",np.ones(n,dtype=np.float32)),("zeros",np.zeros(n,dtype=np.float32)),("half zeros",(np.arange(n)%2).astype(np.float32)),("tiny 1e-30",np.full(n,1e-30,dtype=np.float32))):
    x=fc.RandomVariableCuda(0.0,arr.astype(np.float64))
    ts=[]
    for i in range(5):
        capi.check(capi.load().fmc_sync()); capi.timer_start(); y=x.div(1.1); capi.check(capi.load().fmc_flush()); ts.append(capi.timer_stop())
    print("div(1.1) on",name, "ms", min(ts))
    del x,y
P
