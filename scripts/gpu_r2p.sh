set -x
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_workloads.py -m gpu -x -q -k "brownian or mt19937 or black_scholes or lmm_simulation" 2>&1 | tail -4
python - <<'PY'
import sys, time
sys.path.insert(0, "finmath-lib-cuda-extensions_b200"); sys.path.insert(0, ".")
import finmath_cuda as fc
from finmath_cuda import _capi as capi
fc.ensure_init()
for (T, F, n) in ((80, 1, 1 << 20), (100, 1, 1 << 20), (40, 6, 1 << 20), (80, 1, 100000), (100, 1, 10000000)):
    td = fc.TimeDiscretization(0.0, T, 0.5)
    ts = []
    for i in range(4):
        capi.sync(); t0 = time.perf_counter()
        bm = fc.BrownianMotionCuda(td, F, n, 31415 + (i == 0))
        inc = bm.getBrownianIncrement(0, 0); capi.sync(); ts.append(time.perf_counter() - t0); del bm, inc
    print(f"T={T} F={F} n={n}: first {1e3 * ts[0]:.2f} ms, repeat {1e3 * min(ts[2:]):.3f} ms = {T * F * n / min(ts[2:]) / 1e9:.1f} G increments/s = {4 * T * F * n / min(ts[2:]) / 1e9:.0f} GB/s", flush=True)
PY
