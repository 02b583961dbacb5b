set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout -s KILL 120 python - <<'P' > gpurun_out/first_smoke.log 2>&1
import sys
sys.path[:0]=['.','finmath-lib-cuda-extensions_b200']
import finmath_cuda as fc, numpy as np
fc.ensure_init(0)
n=100000
a=np.random.rand(n); b=np.random.rand(n)
x=fc.RandomVariableCuda(0.0,a); y=fc.RandomVariableCuda(0.0,b)
r=x.add(y).getRealizationsFloat()
print("add ok", np.array_equal(r, a.astype(np.float32)+b.astype(np.float32)))
print("avg", x.mult(y).getAverage(), (a.astype(np.float32)*b.astype(np.float32)).astype(np.float64).mean())
print("exp", np.abs(x.exp().getRealizationsFloat()-np.exp(a.astype(np.float32).astype(np.float64)).astype(np.float32)).max())
P
echo "smoke rc=$?"; cat gpurun_out/first_smoke.log
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python benchmarks/raw_ops.py --sizes 1048576,67108864 --out gpurun_out/raw_ops_r1b.json > gpurun_out/raw_ops_r1b.log 2>&1; echo "raw rc=$?"; tail -5 gpurun_out/raw_ops_r1b.log
timeout -s KILL 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; echo "bench rc=$?"; cat gpurun_out/bench_r1b.json | cut -c1-1500
