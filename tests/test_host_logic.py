"""CPU tests (no GPU): the C-ABI library loads and exports every declared symbol, the host mirror's deterministic /
type-priority logic, loud failure without a device, and the N>1 sharding logic over gloo (world_size 2)."""
import ctypes as C
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fmcuda.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fmc_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    import finmath_cuda
    lib = finmath_cuda._capi.load()
    names = declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fmcuda.h but not exported by libfmcuda.so"
    # the Python binding covers the whole header, nothing more
    assert sorted(finmath_cuda._capi.PROTOTYPES) == names
    # the product library neither links nor contains the oracle
    out = subprocess.run(["nm", "-D", finmath_cuda._capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in out
    ldd = subprocess.run(["ldd", finmath_cuda._capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "torch" not in ldd


def test_driver_library_links_only_the_product():
    from finmath_cuda.workloads import CUDA_DRIVER_LIB
    out = subprocess.run(["nm", "-D", "--undefined-only", CUDA_DRIVER_LIB], capture_output=True, text=True).stdout
    assert "fmc_op_vv" in out and "orc_" not in out


def has_gpu() -> bool:
    import finmath_cuda
    n = C.c_int(0)
    finmath_cuda._capi.load().fmc_device_count(C.byref(n))
    return n.value > 0


def test_no_cpu_fallback_without_device():
    """The product must fail loudly when no CUDA device is present (this container has none)."""
    if has_gpu():
        pytest.skip("a GPU is present")
    import finmath_cuda as fc
    with pytest.raises(fc.CudaError):
        fc.RandomVariableCuda(0.0, [1.0, 2.0, 3.0])
    h = C.c_uint64(0)
    rc = fc._capi.load().fmc_vec_fill(1.0, 10, C.byref(h))
    assert rc == fc._capi.FMC_ERR_NOT_INIT and b"fmc_init" in fc._capi.load().fmc_last_error()


def test_deterministic_random_variables_stay_on_the_host():
    """RandomVariableGPUTest.java:68-86 (testRandomVariableDeterministc): deterministic arithmetic is done in double."""
    import finmath_cuda as fc
    f = fc.RandomVariableCudaFactory()
    rv = f.createRandomVariable(2.0)
    rv = rv.mult(2.0).add(1.0).squared().sub(4.0).div(7.0)
    assert rv.getAverage() == 3.0 and rv.getVariance() == 0.0
    assert rv.isDeterministic() and rv.size() == 1 and rv.getTypePriority() == 20
    a, b = fc.RandomVariableCuda(1.0, 3.0), fc.RandomVariableCuda(2.5, -2.0)
    assert a.add(b).doubleValue() == 1.0 and a.add(b).getFiltrationTime() == 2.5          # time = max (RVF:968)
    assert a.sub(b).doubleValue() == 5.0 and a.bus(b).doubleValue() == -5.0
    assert a.mult(b).doubleValue() == -6.0 and a.div(b).doubleValue() == -1.5 and a.vid(b).doubleValue() == -2.0 / 3.0
    assert a.cap(b).doubleValue() == -2.0 and a.floor(b).doubleValue() == 3.0
    assert a.accrue(b, 0.5).doubleValue() == 3.0 * (1.0 + -2.0 * 0.5) and a.discount(b, 0.25).doubleValue() == 3.0 / (1.0 - 0.5)
    assert a.addProduct(b, 2.0).doubleValue() == -1.0 and a.addProduct(b, b).doubleValue() == 7.0
    assert a.addRatio(a, b).doubleValue() == 1.5 and a.subRatio(a, b).doubleValue() == 4.5
    assert a.pow(2.0).doubleValue() == 9.0 and a.sqrt().doubleValue() == math.sqrt(3.0) and a.invert().doubleValue() == 1 / 3.0
    assert a.exp().doubleValue() == math.exp(3.0) and a.log().doubleValue() == math.log(3.0) and b.abs().doubleValue() == 2.0
    assert a.isNaN().doubleValue() == 0.0 and fc.RandomVariableCuda(float("nan")).isNaN().doubleValue() == 1.0
    assert a.choose(a, b) is a and b.choose(a, b) is b                                   # RVF:1270-1276 returns the object
    assert a.getRealizations().tolist() == [3.0] and a.get(5) == 3.0
    assert a.getMin() == a.getMax() == a.getQuantile(0.3) == 3.0 and a.getStandardError() == 0.0
    assert a.average().doubleValue() == 3.0 and a.getAverage(b) == 3.0 * -2.0
    assert a.div(fc.RandomVariableCuda(0.0)).doubleValue() == math.inf                   # Java double division
    with pytest.raises(NotImplementedError):
        a.apply(lambda v: v)                                                             # RVC:1145-1148


class HigherPriority:
    """A RandomVariable of a type with higher priority (like RandomVariableDifferentiableAAD): it must take over."""

    def __init__(self): self.calls = []
    def getTypePriority(self): return 100
    def getFiltrationTime(self): return 0.0
    def isDeterministic(self): return True
    def doubleValue(self): return 1.0

    def __getattr__(self, name):
        def f(*args):
            self.calls.append(name)
            return self
        return f


def test_type_priority_redispatch():
    """An operand with higher type priority takes over; mirror methods per RVF:962-1178 (vid -> div, fixing RVC:1513-1516)."""
    import finmath_cuda as fc
    import finmath_cuda.random_variable as rvmod
    x = fc.RandomVariableCuda(0.0, 2.0)
    expect = {"add": ["add"], "sub": ["bus"], "bus": ["sub"], "mult": ["mult"], "div": ["vid"], "vid": ["div"], "cap": ["cap"], "floor": ["floor"]}
    for name, want in expect.items():
        h = HigherPriority()
        h.__class__ = type("HP", (HigherPriority, rvmod.RandomVariable), {})
        getattr(x, name)(h)
        assert h.calls == want, (name, h.calls)
    h = HigherPriority(); h.__class__ = type("HP", (HigherPriority, rvmod.RandomVariable), {})
    x.accrue(h, 0.5); assert h.calls == ["mult", "add", "mult"]                          # RVF:1204-1207
    h.calls.clear(); x.discount(h, 0.5); assert h.calls == ["mult", "add", "vid"]        # RVF:1232-1235 (RVC:1606 differs: defect)
    h.calls.clear(); x.addProduct(h, 2.0); assert h.calls == ["mult", "add"]
    h.calls.clear(); x.addRatio(h, x); assert h.calls == ["div", "add"]
    h.calls.clear(); x.subRatio(h, x); assert h.calls == ["div", "mult", "add"]


def test_path_slices_cover_the_range():
    from finmath_cuda.distributed import path_slice, stream_word_offset
    for n in (1, 7, 1000, 1_000_000, 1_048_576):
        for world in (1, 2, 4, 8):
            slices = [path_slice(n, r, world) for r in range(world)]
            assert slices[0][0] == 0 and slices[-1][1] == n
            for (a0, a1), (b0, b1) in zip(slices, slices[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(s[0] % 4 == 0 for s in slices if s[1] > s[0])
    assert stream_word_offset(10, 80, 1) == 1600


WORKER = r'''
import os, sys, json
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "finmath-lib-cuda-extensions_b200"))
from finmath_cuda.distributed import path_slice
from oracle import oracle as O
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
T, F, n, seed = 6, 2, 1001, 31415
p0, p1 = path_slice(n, rank, world)
sq = np.sqrt(np.full(T, 0.5))
mine = O.brownian(seed, T, F, n, sq, p0=p0, p1=p1)          # this rank's slice of every increment vector
# a payoff on the slice, reduced exactly like the runtime does: per-rank (count, sum) partials, summed in rank order
x = O.op_vs(O.FLOOR, O.op_vvs(O.ADDPRODUCT, mine[0], mine[3], 0.3), 0.0)
part = torch.tensor([float(x.size), O.average(x) * x.size if x.size else 0.0], dtype=torch.float64)
gathered = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
dist.all_gather(gathered, part)
cnt = sum(float(g[0]) for g in gathered); tot = sum(float(g[1]) for g in gathered)
# moments merged with Chan's formula in rank order
m = torch.tensor([float(x.size), O.average(x) if x.size else 0.0, O.variance(x) * x.size if x.size > 1 else 0.0], dtype=torch.float64)
gm = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
dist.all_gather(gm, m)
c = v = m2 = 0.0
for g in gm:
    bc, bv, bm = (float(t) for t in g)
    if bc == 0: continue
    if c == 0: c, v, m2 = bc, bv, bm; continue
    tot_c = c + bc; d = bv - v; w = bc / tot_c
    m2 = m2 + bm + d * d * c * w; v = v + d * w; c = tot_c
if rank == 0:
    full = O.brownian(seed, T, F, n, sq)
    xf = O.op_vs(O.FLOOR, O.op_vvs(O.ADDPRODUCT, full[0], full[3], 0.3), 0.0)
    print(json.dumps({{"cnt": cnt, "avg": tot / cnt, "avg_full": O.average(xf), "var": m2 / c, "var_full": O.variance(xf)}}))
dist.destroy_process_group()
'''


def test_world_size_2_sharding_over_gloo(tmp_path):
    """N>1 path: every rank owns a contiguous path slice; the only exchange is a few doubles per reduction."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    res = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert res["cnt"] == 1001
    assert abs(res["avg"] - res["avg_full"]) <= 1e-12 * abs(res["avg_full"])
    assert abs(res["var"] - res["var_full"]) <= 1e-10 * abs(res["var_full"])


def test_java_shim_binds_only_declared_symbols_with_matching_arity():
    """java/.../FmCuda.java cannot be compiled here (no JDK): at least every downcall handle must name a function of
    include/fmcuda.h, with as many parameters as the C prototype has, and the RandomVariable surface must be complete."""
    java_dir = os.path.join(ROOT, "java", "src", "main", "java", "net", "finmath", "cuda", "montecarlo")
    src = open(os.path.join(java_dir, "FmCuda.java")).read()
    header = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    protos = {m.group(1): m.group(2) for m in re.finditer(r"\bint\s+(fmc_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", header, flags=re.S)}
    protos["fmc_last_error"] = "void"
    bound = re.findall(r'h\("(fmc_[a-z0-9_]+)",\s*FunctionDescriptor\.of\(([^;]*?)\)\);', src)
    assert len(bound) >= 24
    for name, desc in bound:
        assert name in protos, f"FmCuda.java binds {name}, which include/fmcuda.h does not declare"
        params = protos[name].strip()
        n_c = 0 if params in ("", "void") else len(params.split(","))
        n_java = len([a for a in desc.split(",") if a.strip()]) - 1          # first layout is the return value
        assert n_c == n_java, (name, n_c, n_java)
    rvc = open(os.path.join(java_dir, "RandomVariableCuda.java")).read()
    for method in ("equals", "getFiltrationTime", "getTypePriority", "get", "size", "getMin", "getMax", "getAverage", "getVariance", "getSampleVariance",
                   "getStandardDeviation", "getStandardError", "getQuantile", "getQuantileExpectation", "getHistogram", "isDeterministic", "cache",
                   "getRealizations", "doubleValue", "getOperator", "getRealizationsStream", "apply", "cap", "floor", "add", "sub", "bus", "mult", "div",
                   "vid", "pow", "average", "squared", "sqrt", "invert", "abs", "exp", "log", "sin", "cos", "accrue", "discount", "choose", "addProduct",
                   "addRatio", "subRatio", "isNaN"):
        assert re.search(r"public\s+\S+(\[\])*\s+%s\(" % method, rvc), f"RandomVariableCuda.java lacks {method}()"
    assert len(re.findall(r"@Override", rvc)) >= 64                          # RandomVariableCuda.java:785-1701 overrides 64 methods
    assert "serialVersionUID = 7620120320663270600L" in rvc and "transient long" in rvc
    assert rvc.count("{") == rvc.count("}") and src.count("{") == src.count("}")


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) on a tiny sample: one JSON line with the
    contract's keys. Runs the oracle only — no GPU needed."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-paths-per-thread", "64",
                        "--steps", "1", "--warmup", "0", "--no-calibration"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "lmm_atm_path_steps_per_s" and line["unit"] == "path-steps/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_host_mirror_objects_recycled_across_valuation_threads():
    """The C++ host mirror (include/finmath/RandomVariableImpl.hpp) takes the object of every recorded operation from a per-thread free
    list; products valued by several host threads free objects on another thread than the one that made them. Run on the oracle twin
    of the driver (same mirror source, CPU backend): the values do not depend on the number of valuation threads, repeatedly."""
    import numpy as np
    from oracle.workloads_oracle import driver
    m = driver().lmm(257)
    ref = np.asarray(m.step()).copy()
    for threads in (3, 1, 4, 2):
        m.set_valuation_threads(threads)
        for _ in range(2):
            assert np.array_equal(np.asarray(m.step()), ref), threads
    m.close()
