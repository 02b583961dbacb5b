"""RandomVariable on the CPU oracle: the RandomVariableFromFloatArray twin as a Python RandomVariable (float32 storage,
every operation through oracle/fm_oracle.c, deterministic values as doubles, type priority 1 — RVF:47).

TEST INFRASTRUCTURE ONLY. It exists so that code written against the RandomVariable interface — in particular the
RandomVariableDifferentiableAAD wrapper of the product package — can be run on the reference semantics and compared with
the same code on RandomVariableCuda (tests/test_gpu_aad.py). Nothing under finmath-lib-cuda-extensions_b200/ imports it.
"""
from __future__ import annotations

import math

import numpy as np

from finmath_cuda.random_variable import RandomVariable      # the interface only (no device code is touched)

from . import oracle as O


class OracleRandomVariable(RandomVariable):
    typePriority = 1                                          # RVF:47

    def __init__(self, *args):
        """OracleRandomVariable(value) | (time, value) | (time, values)."""
        if len(args) == 1:
            time, data = -1.7976931348623157e308, args[0]
        else:
            time, data = args[0], args[1]
        self.time = float(time)
        if np.isscalar(data):
            self.data, self.value = None, float(data)
        else:
            a = np.asarray(data)
            self.data = a.astype(np.float32) if a.dtype == np.float32 else O.from_f64(np.ascontiguousarray(a, dtype=np.float64))   # RVF:217-223
            self.value = math.nan

    # ---- interface ----
    def getTypePriority(self) -> int: return self.typePriority
    def getFiltrationTime(self) -> float: return self.time
    def isDeterministic(self) -> bool: return self.data is None
    def size(self) -> int: return 1 if self.data is None else int(self.data.size)
    def get(self, i: int) -> float: return self.value if self.data is None else float(self.data[i])
    def getRealizations(self) -> np.ndarray: return np.array([self.value]) if self.data is None else self.data.astype(np.float64)
    def doubleValue(self) -> float: return self.value
    def getAverage(self, *a) -> float: return self.value if self.data is None else O.average(self.data)
    def getVariance(self, *a) -> float: return 0.0 if self.data is None else O.variance(self.data)
    def getMin(self) -> float: return self.value if self.data is None else float(O.minimum(self.data))
    def getMax(self) -> float: return self.value if self.data is None else float(O.maximum(self.data))

    # ---- helpers ----
    def _t(self, other) -> float:
        return max(self.time, other.time) if isinstance(other, OracleRandomVariable) else self.time

    def _takes_over(self, other) -> bool:
        return isinstance(other, RandomVariable) and not isinstance(other, OracleRandomVariable) and other.getTypePriority() > self.typePriority

    def _bin(self, code, mirror, dfun, other):
        """this (op) other with RVF's dispatch: priority hand-over, deterministic pairs in double, scalars cast to float."""
        if self._takes_over(other):
            return getattr(other, mirror)(self)                                               # RVF:962-965 pattern
        if not isinstance(other, RandomVariable):
            other = OracleRandomVariable(float(other))
        t = self._t(other)
        if self.data is None and other.data is None:
            return OracleRandomVariable(t, dfun(self.value, other.value))                     # RVF:970-973: double
        if other.data is None:
            return OracleRandomVariable(t, O.op_vs(code, self.data, other.value))
        if self.data is None:
            rev = {O.ADD: O.ADD, O.MULT: O.MULT, O.SUB: O.BUS, O.BUS: O.SUB, O.DIV: O.VID, O.VID: O.DIV, O.CAP: O.CAP, O.FLOOR: O.FLOOR}[code]
            return OracleRandomVariable(t, O.op_vs(rev, other.data, self.value))
        return OracleRandomVariable(t, O.op_vv(code, self.data, other.data))

    def _un(self, code, dfun):
        if self.data is None:
            return OracleRandomVariable(self.time, dfun(self.value))
        return OracleRandomVariable(self.time, O.op_v(code, self.data))

    # ---- operations ----
    def add(self, x): return self._bin(O.ADD, "add", lambda a, b: a + b, x)
    def sub(self, x): return self._bin(O.SUB, "bus", lambda a, b: a - b, x)
    def bus(self, x): return self._bin(O.BUS, "sub", lambda a, b: b - a, x)
    def mult(self, x): return self._bin(O.MULT, "mult", lambda a, b: a * b, x)
    def div(self, x): return self._bin(O.DIV, "vid", lambda a, b: a / b if b != 0 else math.copysign(math.inf, a) if a == a and a != 0 else math.nan, x)
    def vid(self, x): return self._bin(O.VID, "div", lambda a, b: b / a if a != 0 else math.copysign(math.inf, b) if b == b and b != 0 else math.nan, x)
    def cap(self, x): return self._bin(O.CAP, "cap", min, x)
    def floor(self, x): return self._bin(O.FLOOR, "floor", max, x)
    def squared(self): return self._un(O.SQUARED, lambda a: a * a)
    def sqrt(self): return self._un(O.SQRT, lambda a: math.sqrt(a) if a >= 0 else math.nan)
    def exp(self): return self._un(O.EXP, math.exp)
    def log(self): return self._un(O.LOG, lambda a: math.log(a) if a > 0 else (-math.inf if a == 0 else math.nan))
    def invert(self): return self._un(O.INVERT, lambda a: 1.0 / a if a != 0 else math.inf)
    def abs(self): return self._un(O.ABS, abs)

    def pow(self, e: float):
        if self.data is None:
            return OracleRandomVariable(self.time, math.pow(self.value, e))
        return OracleRandomVariable(self.time, O.op_vs(O.POW, self.data, float(e)))

    def average(self): return OracleRandomVariable(self.time, self.getAverage())

    def _vec(self, n):
        return self.data if self.data is not None else np.full(n, np.float32(self.value), dtype=np.float32)

    def accrue(self, rate, p: float):
        if self._takes_over(rate): return rate.mult(p).add(1.0).mult(self)                  # RVF:1204-1207
        if not isinstance(rate, RandomVariable): rate = OracleRandomVariable(float(rate))
        if self.data is None and rate.data is None: return OracleRandomVariable(self._t(rate), self.value * (1.0 + rate.value * p))
        n = max(self.size(), rate.size())
        return OracleRandomVariable(self._t(rate), O.op_vvs(O.ACCRUE, self._vec(n), rate._vec(n), float(p)))

    def discount(self, rate, p: float):
        if self._takes_over(rate): return rate.mult(p).add(1.0).vid(self)                   # RVF:1232-1235
        if not isinstance(rate, RandomVariable): rate = OracleRandomVariable(float(rate))
        if self.data is None and rate.data is None: return OracleRandomVariable(self._t(rate), self.value / (1.0 + rate.value * p))
        n = max(self.size(), rate.size())
        return OracleRandomVariable(self._t(rate), O.op_vvs(O.DISCOUNT, self._vec(n), rate._vec(n), float(p)))

    def addProduct(self, f1, f2):
        if self._takes_over(f1): return f1.mult(f2).add(self)                               # RVF:1319-1322
        if isinstance(f2, RandomVariable) and self._takes_over(f2): return f2.mult(f1).add(self)
        if not isinstance(f1, RandomVariable): f1 = OracleRandomVariable(float(f1))
        if isinstance(f2, RandomVariable):
            if self.data is None and f1.data is None and f2.data is None:
                return OracleRandomVariable(max(self._t(f1), f2.time), self.value + f1.value * f2.value)
            n = max(self.size(), f1.size(), f2.size())
            return OracleRandomVariable(max(self._t(f1), f2.time), O.op_vvv(O.ADDPRODUCT, self._vec(n), f1._vec(n), f2._vec(n)))
        if self.data is None and f1.data is None:
            return OracleRandomVariable(self._t(f1), self.value + f1.value * float(f2))
        n = max(self.size(), f1.size())
        return OracleRandomVariable(self._t(f1), O.op_vvs(O.ADDPRODUCT, self._vec(n), f1._vec(n), float(f2)))

    def choose(self, a, b):
        if not isinstance(a, RandomVariable): a = OracleRandomVariable(float(a))
        if not isinstance(b, RandomVariable): b = OracleRandomVariable(float(b))
        if self.data is None:
            return a if self.value >= 0 else b                                              # RVF:1264-1270
        n = self.size()
        return OracleRandomVariable(max(self.time, a.time, b.time), O.op_vvv(O.CHOOSE, self.data, a._vec(n), b._vec(n)))
