"""Kernel-level cross-check against the UNMODIFIED reference kernels: RandomVariableCudaKernel.cu compiled for sm_100a
with the reference's own flags (`-fmad false`, oracle/Makefile target `ref`, output oracle/_ref/) is launched through
the driver API with the reference's geometry on the same inputs as the product's fused interpreter.
Exact ops must agree bit for bit. exp/log/pow are expf/logf/powf in the reference's GPU kernels
(RandomVariableCudaKernel.cu:103,124,134) but double-then-round in its CPU float class (the specification the product
follows, RandomVariableFromFloatArray.java:849,890,905): there the two may differ by a couple of ulps."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 200_003


@pytest.fixture(scope="module")
def ref(fc):
    from oracle import ref_kernels
    if not ref_kernels.available():
        pytest.skip("oracle/_ref cubin or cuda-python not available")
    dev = ctypes.c_double()
    fc._capi.check(fc._capi.load().fmc_get_option(b"device_index", ctypes.byref(dev)))     # the runtime's device (-1 = last by default)
    return ref_kernels.ReferenceKernels(int(dev.value))


def dev_ptr(fc, rv) -> int:
    p = ctypes.c_void_p()
    fc._capi.check(fc._capi.load().fmc_vec_device_ptr(rv.handle, ctypes.byref(p)))
    return p.value


def run_ref(fc, ref, name, n, ins, scalars=()):
    out = np.empty(n, dtype=np.float32)
    h = ctypes.c_uint64()
    L = fc._capi.load()
    fc._capi.check(L.fmc_vec_alloc(n, ctypes.byref(h)))
    p = ctypes.c_void_p()
    fc._capi.check(L.fmc_vec_device_ptr(h.value, ctypes.byref(p)))
    fc._capi.check(L.fmc_sync())
    args = [("p", dev_ptr(fc, v)) for v in ins] + [("f", s) for s in scalars] + [("p", p.value)]
    ref.launch(name, n, args)
    ref.synchronize()
    fc._capi.check(L.fmc_vec_to_f32(h.value, out.ctypes.data, n))
    fc._capi.check(L.fmc_vec_release(h.value))
    return out


def ulps(a, b):
    ia = a.view(np.int32).astype(np.int64); ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7fffffff), ia); ib = np.where(ib < 0, -(ib & 0x7fffffff), ib)
    return int(np.abs(ia - ib).max())


@pytest.fixture(scope="module")
def vecs(fc):
    rng = np.random.default_rng(2024)
    a = (rng.random(N, dtype=np.float32) * 4 - 2).astype(np.float32)
    b = (rng.random(N, dtype=np.float32) + np.float32(0.25)).astype(np.float32)
    c = (rng.random(N, dtype=np.float32) - np.float32(0.5)).astype(np.float32)
    A, B, C_ = (fc.RandomVariableCuda(0.0, v.astype(np.float64)) for v in (a, b, c))
    return A, B, C_


@pytest.mark.parametrize("kernel,method,s", [
    ("capByScalar", "cap", 0.3), ("floorByScalar", "floor", -0.1), ("addScalar", "add", 3.1415), ("subScalar", "sub", 1.0 / 3.0),
    ("busScalar", "bus", 2.0), ("multScalar", "mult", 1.0 / 3.0), ("divScalar", "div", 3.1415), ("vidScalar", "vid", 2.0)])
def test_scalar_kernels_bit_exact(fc, ref, vecs, kernel, method, s):
    A, _, _ = vecs
    want = run_ref(fc, ref, kernel, N, [A], [s])
    got = getattr(A, method)(s).getRealizationsFloat()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("kernel,method", [("add", "add"), ("sub", "sub"), ("mult", "mult"), ("cuDiv", "div"), ("cap", "cap"), ("cuFloor", "floor")])
def test_vector_kernels_bit_exact(fc, ref, vecs, kernel, method):
    A, B, _ = vecs
    want = run_ref(fc, ref, kernel, N, [A, B])
    got = getattr(A, method)(B).getRealizationsFloat()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("kernel,method", [("cuSqrt", "sqrt"), ("invert", "invert"), ("cuAbs", "abs"), ("squared", "squared")])
def test_unary_kernels_bit_exact(fc, ref, vecs, kernel, method):
    _, B, _ = vecs
    want = run_ref(fc, ref, kernel, N, [B])
    got = getattr(B, method)().getRealizationsFloat()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_compound_kernels_bit_exact(fc, ref, vecs):
    A, B, C_ = vecs
    for kernel, got in (("accrue", A.accrue(B, 0.5)), ("discount", A.discount(B, 0.5))):
        want = run_ref(fc, ref, kernel, N, [A, B], [0.5])
        assert np.array_equal(got.getRealizationsFloat().view(np.uint32), want.view(np.uint32)), kernel
    want = run_ref(fc, ref, "addProduct_vs", N, [A, B], [0.3])
    assert np.array_equal(A.addProduct(B, 0.3).getRealizationsFloat().view(np.uint32), want.view(np.uint32))
    want = run_ref(fc, ref, "addProduct", N, [A, B, C_])
    assert np.array_equal(A.addProduct(B, C_).getRealizationsFloat().view(np.uint32), want.view(np.uint32))
    want = run_ref(fc, ref, "addRatio", N, [A, C_, B])
    assert np.array_equal(A.addRatio(C_, B).getRealizationsFloat().view(np.uint32), want.view(np.uint32))
    want = run_ref(fc, ref, "subRatio", N, [A, C_, B])
    assert np.array_equal(A.subRatio(C_, B).getRealizationsFloat().view(np.uint32), want.view(np.uint32))


def test_transcendental_kernels_close(fc, ref, vecs):
    _, B, _ = vecs
    for kernel, got in (("cuExp", B.exp()), ("cuLog", B.log())):
        want = run_ref(fc, ref, kernel, N, [B])
        assert ulps(got.getRealizationsFloat(), want) <= 2, kernel          # expf/logf vs double-then-round
    want = run_ref(fc, ref, "cuPow", N, [B], [1.7])
    assert ulps(B.pow(1.7).getRealizationsFloat(), want) <= 4               # powf is documented to a few ulp
