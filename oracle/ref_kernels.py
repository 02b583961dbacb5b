"""TEST INFRASTRUCTURE: launches the REFERENCE's own CUDA kernels (RandomVariableCudaKernel.cu, compiled where it lies by
`make -C oracle ref` into oracle/_ref/, never copied into this repository) through the CUDA driver API, with the
reference's launch geometry (1024 threads per block, one element per thread, default stream — RandomVariableCuda.java:539-557).

Two uses: (1) a kernel-level cross-check of the product's elementwise results against the unmodified reference kernels on
the same GPU; (2) the "reference kernel on B200" bandwidth bar in benchmarks/raw_ops.py. Only tests/ and benchmarks/ import this.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CUBIN = os.path.join(_HERE, "_ref", "RandomVariableCudaKernel.sm_100a.cubin")


def available() -> bool:
    if not os.path.exists(CUBIN):
        return False
    try:
        from cuda.bindings import driver  # noqa: F401
        return True
    except Exception:
        return False


class ReferenceKernels:
    BLOCK = 1024                                                   # RandomVariableCuda.java:545

    def __init__(self, device_index: int = 0):
        from cuda.bindings import driver as drv
        self.drv = drv
        self._ck(drv.cuInit(0))
        dev = self._ck(drv.cuDeviceGet(device_index))
        self.ctx = self._ck(drv.cuDevicePrimaryCtxRetain(dev))      # share the runtime API's primary context
        self._ck(drv.cuCtxSetCurrent(self.ctx))
        with open(CUBIN, "rb") as f:
            image = f.read()
        self.module = self._ck(drv.cuModuleLoadData(image))
        self.funcs = {}

    @staticmethod
    def _ck(res):
        err = res[0]
        if int(err) != 0:
            raise RuntimeError(f"CUDA driver error {err}")
        return res[1] if len(res) == 2 else res[1:] if len(res) > 2 else None

    def _fn(self, name: str):
        if name not in self.funcs:
            self.funcs[name] = self._ck(self.drv.cuModuleGetFunction(self.module, name.encode()))
        return self.funcs[name]

    def launch(self, name: str, n: int, args: list) -> None:
        """args: ('p', device_ptr) | ('f', float) in the kernel's order after the leading `int n`."""
        vals = [ctypes.c_int(n)]
        for kind, v in args:
            vals.append(ctypes.c_void_p(int(v)) if kind == "p" else ctypes.c_float(float(v)))
        ptrs = (ctypes.c_void_p * len(vals))(*[ctypes.cast(ctypes.pointer(v), ctypes.c_void_p) for v in vals])
        grid = max(1, (n + self.BLOCK - 1) // self.BLOCK)
        self._ck(self.drv.cuLaunchKernel(self._fn(name), grid, 1, 1, self.BLOCK, 1, 1, 0, 0, ctypes.addressof(ptrs), 0))

    def synchronize(self) -> None:
        self._ck(self.drv.cuCtxSynchronize())

    def time_ms(self, name: str, n: int, args: list, repeats: int = 5) -> float:
        drv = self.drv
        e0 = self._ck(drv.cuEventCreate(0)); e1 = self._ck(drv.cuEventCreate(0))
        best = None
        for _ in range(repeats + 2):
            self._ck(drv.cuEventRecord(e0, 0))
            self.launch(name, n, args)
            self._ck(drv.cuEventRecord(e1, 0))
            self._ck(drv.cuEventSynchronize(e1))
            ms = self._ck(drv.cuEventElapsedTime(e0, e1))
            best = ms if best is None else min(best, ms)
        return float(best)
