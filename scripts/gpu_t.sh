timeout -s KILL 600 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout -s KILL 120 python - <<'P'
import sys, time
sys.path[:0]=['.','finmath-lib-cuda-extensions_b200']
import finmath_cuda as fc, numpy as np
fc.ensure_init()
for n in (1_000_000, 100_000_000):
    x = fc.RandomVariableCuda(0.0, np.random.default_rng(1).standard_normal(n).astype(np.float32))
    for name, f in (("getQuantile(0.95)", lambda: x.getQuantile(0.95)), ("getQuantileExpectation(0.05,0.95)", lambda: x.getQuantileExpectation(0.05, 0.95)), ("getHistogram(9 points)", lambda: x.getHistogram(np.linspace(-2, 2, 9)))):
        f(); t0 = time.perf_counter(); r = f(); dt = time.perf_counter() - t0
        print(f"n={n:>10} {name:36s} {dt * 1e3:8.3f} ms")
P
