import sys, time
sys.path[:0]=['.','finmath-lib-cuda-extensions_b200']
import finmath_cuda as fc
from finmath_cuda.workloads import DriverLib
from oracle.workloads_oracle import driver
fc.ensure_init()
gpu, cpu = DriverLib(), driver()
spec = (10, 30, 2, 40, 0.02)
for n in (20000, 100000, 300000, 1000000):
    mg = gpu.lmm(n); mg.simulate(); vg = mg.bermudan(*spec); vg2 = mg.bermudan(*spec)
    if n <= 300000:
        mc = cpu.lmm(n); mc.simulate(); vc = mc.bermudan(*spec)
    else:
        vc = float('nan')
    print(n, "gpu", vg, vg2, "cpu", vc, flush=True)
