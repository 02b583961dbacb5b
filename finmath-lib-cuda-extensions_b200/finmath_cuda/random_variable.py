"""RandomVariableCuda / RandomVariableCudaFactory — host-side mirror of the reference's plug-in surface.

Mirrors, method for method, net.finmath.cuda.montecarlo.RandomVariableCuda
(/root/reference/src/main/java/net/finmath/cuda/montecarlo/RandomVariableCuda.java, "RVC") with the semantics of its
CPU twin RandomVariableFromFloatArray ("RVF", .../cuda/cpu/montecarlo/RandomVariableFromFloatArray.java), which is
the behavioural specification wherever RVC is incomplete or defective:
  * data model: stochastic = fp32 device vector (an fmc_vec handle) + filtration time; deterministic = one Python
    float (double), size() == 1 (RVC:566-577, 821-827); type priority 20 (RVC:568);
  * deterministic (op) deterministic is computed in double on the host (RVC:1400-1403); a deterministic operand of
    a stochastic vector enters the device op as a scalar cast to float (RVC:521);
  * an operand with a higher type priority takes over the operation (RVC:1392-1395, mirror methods per RVF);
  * choose / isNaN / sin / cos, which RVC leaves unimplemented (RVC:1632-1635, 1701-1704, 1356, 1372), follow RVF.
All device work goes through the C ABI (include/fmcuda.h); operations are recorded and fused, nothing here computes
on the CPU and there is no fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, Sequence

import numpy as np

from . import _capi as capi


class RandomVariable:
    """The subset of net.finmath.stochastic.RandomVariable that the hot path uses (marker base class)."""

    def getTypePriority(self) -> int: raise NotImplementedError
    def getFiltrationTime(self) -> float: raise NotImplementedError
    def isDeterministic(self) -> bool: raise NotImplementedError
    def size(self) -> int: raise NotImplementedError
    def get(self, i: int) -> float: raise NotImplementedError
    def getRealizations(self) -> np.ndarray: raise NotImplementedError
    def doubleValue(self) -> float: raise NotImplementedError


def _java_min(a: float, b: float) -> float:
    if a != a or b != b:
        return math.nan
    if a == 0.0 and b == 0.0:
        return -0.0 if (math.copysign(1.0, a) < 0 or math.copysign(1.0, b) < 0) else 0.0
    return a if a < b else b


def _java_max(a: float, b: float) -> float:
    if a != a or b != b:
        return math.nan
    if a == 0.0 and b == 0.0:
        return -0.0 if (math.copysign(1.0, a) < 0 and math.copysign(1.0, b) < 0) else 0.0
    return a if a > b else b


def _java_pow(x: float, y: float) -> float:
    if y != y:
        return math.nan
    if y == 0.0:
        return 1.0
    if x != x:
        return math.nan
    if math.isinf(y) and abs(x) == 1.0:
        return math.nan
    try:
        return math.pow(x, y)
    except OverflowError:
        return math.inf if (x > 0 or float(y).is_integer() and int(y) % 2 == 0) else -math.inf
    except ValueError:
        if x == 0.0:
            return math.inf
        return math.nan


def _div(a: float, b: float) -> float:
    """Java double division (no ZeroDivisionError)."""
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0.0:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)


class RandomVariableCuda(RandomVariable):
    typePriorityDefault = 20                                    # RVC:568

    __slots__ = ("time", "_h", "_size", "valueIfNonStochastic", "typePriority")

    # ------------------------------------------------------------------ construction (RVC:618-734)
    def __init__(self, *args):
        """RandomVariableCuda(value) | (time, value) | (time, values) | (time, value|values, typePriority)."""
        self._h = 0
        self._size = 1
        self.valueIfNonStochastic = math.nan
        self.typePriority = self.typePriorityDefault
        if len(args) == 1:
            if isinstance(args[0], RandomVariable):               # RVF:64-70 copy constructor
                rv = args[0]
                self.time = rv.getFiltrationTime()
                if rv.isDeterministic():
                    self.valueIfNonStochastic = float(rv.doubleValue())
                else:
                    self._upload(rv.getRealizations())
                return
            time, data = -math.inf if np.isscalar(args[0]) else 0.0, args[0]   # RVC:667-669 uses -Double.MAX_VALUE; RVC:732-734
            if np.isscalar(args[0]):
                time = -1.7976931348623157e308
        elif len(args) == 2:
            time, data = args
        elif len(args) == 3:
            time, data, prio = args
            self.typePriority = int(prio)
        else:
            raise TypeError("RandomVariableCuda(value) | (time, value) | (time, values) | (time, value|values, typePriority)")
        self.time = float(time)
        if np.isscalar(data):
            self.valueIfNonStochastic = float(data)                # RVC:678-684
        else:
            self._upload(data)                                     # RVC:693-695, 723-725

    def _upload(self, data) -> None:
        capi.ensure_init()
        arr = np.asarray(data)
        h = C.c_uint64(0)
        if arr.dtype == np.float32:
            arr = np.ascontiguousarray(arr)
            capi.check(capi.load().fmc_vec_from_f32(arr.ctypes.data, arr.size, C.byref(h)))
        else:
            arr = np.ascontiguousarray(arr, dtype=np.float64)      # (float) cast happens in the runtime, RVC:768-774
            capi.check(capi.load().fmc_vec_from_f64(arr.ctypes.data, arr.size, C.byref(h)))
        self._h = h.value
        self._size = int(arr.size)

    @classmethod
    def of(cls, time: float, value_or_handle, size: int | None = None, typePriority: int | None = None) -> "RandomVariableCuda":
        """RVC:618-646. of(time, value) -> constant; of(time, handle, size) wraps a device vector handle (takes ownership)."""
        rv = cls.__new__(cls)
        rv.time = float(time)
        rv.typePriority = cls.typePriorityDefault if typePriority is None else int(typePriority)
        if size is None:
            rv._h = 0; rv._size = 1; rv.valueIfNonStochastic = float(value_or_handle)
        else:
            rv._h = int(value_or_handle); rv._size = int(size); rv.valueIfNonStochastic = math.nan
        return rv

    def __del__(self):
        h = getattr(self, "_h", 0)
        if h:
            try:
                lib = capi._lib
                if lib is not None and lib.fmc_is_initialized():
                    lib.fmc_vec_release(h)
            except Exception:
                pass
            self._h = 0

    @staticmethod
    def getDevicePointer(size: int) -> int:
        """RVC:737-739: an uninitialised device vector of `size` floats; returns the handle."""
        capi.ensure_init()
        h = C.c_uint64(0)
        capi.check(capi.load().fmc_vec_alloc(int(size), C.byref(h)))
        return h.value

    @staticmethod
    def clean() -> None:                                           # RVC:751-753
        if capi._lib is not None and capi._lib.fmc_is_initialized():
            capi.check(capi.load().fmc_pool_trim())

    @staticmethod
    def purge() -> None:                                           # RVC:755-757
        if capi._lib is not None and capi._lib.fmc_is_initialized():
            capi.check(capi.load().fmc_pool_purge())

    # ------------------------------------------------------------------ helpers
    @property
    def handle(self) -> int:
        return self._h

    @staticmethod
    def _wrap(time: float, h: C.c_uint64, size: int) -> "RandomVariableCuda":
        return RandomVariableCuda.of(time, h.value, size)

    @staticmethod
    def _as_cuda(rv: RandomVariable) -> "RandomVariableCuda":
        """getRandomVariableCuda (RVC:759-766): foreign vectors are uploaded on the fly."""
        if isinstance(rv, RandomVariableCuda):
            return rv
        if rv.isDeterministic():
            return RandomVariableCuda.of(rv.getFiltrationTime(), rv.doubleValue())
        return RandomVariableCuda(rv.getFiltrationTime(), rv.getRealizations())

    def _vs(self, op: int, s: float, time: float | None = None) -> "RandomVariableCuda":
        out = C.c_uint64(0)
        capi.check(capi.load().fmc_op_vs(op, self._h, float(s), C.byref(out)))
        return self._wrap(self.time if time is None else time, out, self._size)

    def _v(self, op: int) -> "RandomVariableCuda":
        out = C.c_uint64(0)
        capi.check(capi.load().fmc_op_v(op, self._h, C.byref(out)))
        return self._wrap(self.time, out, self._size)

    def _vv(self, op: int, other: "RandomVariableCuda", time: float) -> "RandomVariableCuda":
        out = C.c_uint64(0)
        capi.check(capi.load().fmc_op_vv(op, self._h, other._h, C.byref(out)))
        return self._wrap(time, out, self._size)

    def _reduce(self, kind: int, weights: "RandomVariableCuda | None" = None) -> float:
        out = C.c_double(0.0)
        capi.check(capi.load().fmc_reduce(kind, self._h, weights._h if weights is not None else 0, C.byref(out)))
        return out.value

    # ------------------------------------------------------------------ basic accessors
    def equals(self, rv: RandomVariable) -> bool:                  # RVF:233-253 (RVC:785-786 throws)
        if self.time != rv.getFiltrationTime():
            return False
        if self.isDeterministic() and rv.isDeterministic():
            return self.valueIfNonStochastic == rv.doubleValue()
        if self.isDeterministic() != rv.isDeterministic():
            return False
        a, b = self.getRealizations(), rv.getRealizations()
        return a.shape == b.shape and bool(np.all(a == b))

    def getFiltrationTime(self) -> float: return self.time         # RVC:801-804
    def getTypePriority(self) -> int: return self.typePriority     # RVC:806-809
    def isDeterministic(self) -> bool: return self._h == 0         # RVC:1093-1096
    def size(self) -> int: return 1 if self._h == 0 else self._size  # RVC:820-827

    def get(self, pathOrState: int) -> float:                      # RVF:265-272 (RVC:811-818 throws when stochastic)
        if self.isDeterministic():
            return self.valueIfNonStochastic
        out = C.c_double(0.0)
        capi.check(capi.load().fmc_vec_get(self._h, int(pathOrState), C.byref(out)))
        return out.value

    def doubleValue(self) -> float:                                # RVC:1124-1131
        if self.isDeterministic():
            return self.valueIfNonStochastic
        raise NotImplementedError("The random variable is non-deterministic")

    def cache(self) -> "RandomVariableCuda": return self           # RVC:1098-1100

    def getRealizations(self) -> np.ndarray:                       # RVC:1114-1122
        if self.isDeterministic():
            return np.array([self.valueIfNonStochastic], dtype=np.float64)
        out = np.empty(self._size, dtype=np.float64)
        capi.check(capi.load().fmc_vec_to_f64(self._h, out.ctypes.data, out.size))
        return out

    def getRealizationsFloat(self) -> np.ndarray:                  # getValuesAsFloat RVC:469-481
        if self.isDeterministic():
            return np.array([self.valueIfNonStochastic], dtype=np.float32)
        out = np.empty(self._size, dtype=np.float32)
        capi.check(capi.load().fmc_vec_to_f32(self._h, out.ctypes.data, out.size))
        return out

    def getRealizationsStream(self):                               # RVF:614-626 (RVC:1139-1143 returns null)
        return iter(self.getRealizations())

    def getOperator(self) -> Callable[[int], float]:               # RVF:647-664 (RVC:1133-1137 returns null)
        if self.isDeterministic():
            v = self.valueIfNonStochastic
            return lambda i: v
        r = self.getRealizations()
        return lambda i: float(r[i])

    def apply(self, *args):                                        # RVC:1145-1169 throws UnsupportedOperationException
        raise NotImplementedError("apply(function) cannot run on the device (RandomVariableCuda.java:1147)")

    # ------------------------------------------------------------------ statistics
    def getMin(self) -> float:                                     # RVF:283-296
        return self.valueIfNonStochastic if self.isDeterministic() else self._reduce(capi.RED_MIN)

    def getMax(self) -> float:                                     # RVF:298-311
        return self.valueIfNonStochastic if self.isDeterministic() else self._reduce(capi.RED_MAX)

    def getAverage(self, probabilities: RandomVariable | None = None) -> float:
        if probabilities is None:                                  # RVF:313-334
            return self.valueIfNonStochastic if self.isDeterministic() else self._reduce(capi.RED_AVERAGE)
        if self.isDeterministic():                                 # RVF:338-340
            return self.valueIfNonStochastic * probabilities.getAverage()
        p = self._as_cuda(probabilities)
        if p.isDeterministic():
            return self._vs(capi.MULT, p.valueIfNonStochastic).getAverage()
        return self._reduce(capi.RED_AVERAGE_W, p)                 # RVF:341-356

    def getVariance(self, probabilities: RandomVariable | None = None) -> float:
        if probabilities is None:                                  # RVF:359-382
            if self.isDeterministic():
                return 0.0
            return self._reduce(capi.RED_VARIANCE)
        if self.isDeterministic():                                 # RVF:386-388
            return 0.0
        p = self._as_cuda(probabilities)
        if p.isDeterministic():
            p = RandomVariableCuda(p.time, np.full(self._size, p.valueIfNonStochastic))
        return self._reduce(capi.RED_VARIANCE_W, p)                # RVF:389-406

    def getSampleVariance(self) -> float:                          # RVF:409-419
        if self.isDeterministic():
            return 0.0
        return self._reduce(capi.RED_SAMPLE_VARIANCE)

    def getStandardDeviation(self, probabilities: RandomVariable | None = None) -> float:   # RVF:421-443
        if self.isDeterministic():
            return 0.0
        return math.sqrt(self.getVariance(probabilities))

    def getStandardError(self, probabilities: RandomVariable | None = None) -> float:       # RVF:445-470
        if self.isDeterministic():
            return 0.0
        n = self.size()
        if n == 0:
            return math.nan
        return self.getStandardDeviation(probabilities) / math.sqrt(n)

    def getQuantile(self, quantile: float, probabilities: RandomVariable | None = None) -> float:   # RVF:472-499
        if self.isDeterministic():
            return self.valueIfNonStochastic
        if probabilities is not None:
            raise RuntimeError("Method not implemented.")
        out = C.c_double(0.0)
        capi.check(capi.load().fmc_quantile(self._h, float(quantile), C.byref(out)))
        return out.value

    def getQuantileExpectation(self, quantileStart: float, quantileEnd: float) -> float:            # RVF:501-526
        if self.isDeterministic():
            return self.valueIfNonStochastic
        out = C.c_double(0.0)
        capi.check(capi.load().fmc_quantile_expectation(self._h, float(quantileStart), float(quantileEnd), C.byref(out)))
        return out.value

    def getHistogram(self, a, standardDeviations: float | None = None):
        if standardDeviations is not None:                         # RVF:583-602
            numberOfPoints = int(a)
            center = self.getAverage()
            radius = standardDeviations * self.getStandardDeviation()
            stepSize = (numberOfPoints - 1) / 2.0
            intervalPoints = np.empty(numberOfPoints); anchorPoints = np.empty(numberOfPoints + 1)
            for i in range(numberOfPoints):
                alpha = (-(numberOfPoints - 1) / 2.0 + i) / stepSize
                intervalPoints[i] = center + alpha * radius
                anchorPoints[i] = center + alpha * radius - radius / (2 * stepSize)
            anchorPoints[numberOfPoints] = center + 1 * radius + radius / (2 * stepSize)
            return [anchorPoints, self.getHistogram(intervalPoints)]
        pts = np.ascontiguousarray(a, dtype=np.float64)            # RVF:528-581
        out = np.zeros(pts.size + 1, dtype=np.float64)
        if self.isDeterministic():
            for k in range(pts.size):
                if self.valueIfNonStochastic > pts[k]:
                    out[k] = 1.0
                    break
            out[pts.size] = 1.0
            return out
        capi.check(capi.load().fmc_histogram(self._h, pts.ctypes.data, pts.size, out.ctypes.data))
        return out

    # ------------------------------------------------------------------ unary / scalar operators (RVC:1171-1352)
    def _det(self, value: float, time: float | None = None) -> "RandomVariableCuda":
        return RandomVariableCuda.of(self.time if time is None else time, value, None, self.typePriority)

    def cap(self, cap):
        if isinstance(cap, RandomVariable): return self._cap_rv(cap)
        if self.isDeterministic(): return self._det(_java_min(self.valueIfNonStochastic, float(cap)))
        return self._vs(capi.CAP, cap)

    def floor(self, floor):
        if isinstance(floor, RandomVariable): return self._floor_rv(floor)
        if self.isDeterministic(): return self._det(_java_max(self.valueIfNonStochastic, float(floor)))
        return self._vs(capi.FLOOR, floor)

    def add(self, value):
        if isinstance(value, RandomVariable): return self._add_rv(value)
        if self.isDeterministic(): return self._det(self.valueIfNonStochastic + float(value))
        return self._vs(capi.ADD, value)

    def sub(self, value):
        if isinstance(value, RandomVariable): return self._sub_rv(value)
        if self.isDeterministic(): return self._det(self.valueIfNonStochastic - float(value))
        return self._vs(capi.SUB, value)

    def bus(self, value):
        if isinstance(value, RandomVariable): return self._bus_rv(value)
        if self.isDeterministic(): return self._det(-self.valueIfNonStochastic + float(value))
        return self._vs(capi.BUS, value)

    def mult(self, value):
        if isinstance(value, RandomVariable): return self._mult_rv(value)
        if self.isDeterministic(): return self._det(self.valueIfNonStochastic * float(value))
        return self._vs(capi.MULT, value)

    def div(self, value):
        if isinstance(value, RandomVariable): return self._div_rv(value)
        if self.isDeterministic(): return self._det(_div(self.valueIfNonStochastic, float(value)))
        return self._vs(capi.DIV, value)

    def vid(self, value):
        if isinstance(value, RandomVariable): return self._vid_rv(value)
        if self.isDeterministic(): return self._det(_div(float(value), self.valueIfNonStochastic))
        return self._vs(capi.VID, value)

    def pow(self, exponent: float):
        if self.isDeterministic(): return self._det(_java_pow(self.valueIfNonStochastic, float(exponent)))
        return self._vs(capi.POW, exponent)

    def average(self):                                             # RVC:1279-1282
        return RandomVariableCuda.of(-1.7976931348623157e308, self.getAverage())

    def getConditionalExpectation(self, conditionalExpectationOperator):   # RVF:860-864
        return conditionalExpectationOperator.getConditionalExpectation(self)

    def squared(self):
        if self.isDeterministic(): return self._det(self.valueIfNonStochastic * self.valueIfNonStochastic)
        return self._v(capi.SQUARED)

    def sqrt(self):
        if self.isDeterministic():
            v = self.valueIfNonStochastic
            return self._det(math.sqrt(v) if v >= 0 else math.nan)
        return self._v(capi.SQRT)

    def exp(self):
        if self.isDeterministic():
            try: return self._det(math.exp(self.valueIfNonStochastic))
            except OverflowError: return self._det(math.inf)
        return self._v(capi.EXP)

    def log(self):
        if self.isDeterministic():
            v = self.valueIfNonStochastic
            return self._det(math.log(v) if v > 0 else (-math.inf if v == 0 else math.nan))
        return self._v(capi.LOG)

    def sin(self):                                                 # RVF:926-939
        if self.isDeterministic(): return self._det(math.sin(self.valueIfNonStochastic))
        return self._v(capi.SIN)

    def cos(self):                                                 # RVF:941-954
        if self.isDeterministic(): return self._det(math.cos(self.valueIfNonStochastic))
        return self._v(capi.COS)

    def invert(self):
        if self.isDeterministic(): return self._det(_div(1.0, self.valueIfNonStochastic))
        return self._v(capi.INVERT)

    def abs(self):
        if self.isDeterministic(): return self._det(abs(self.valueIfNonStochastic))
        return self._v(capi.ABS)

    def isNaN(self):                                               # RVF:1440-1451 (RVC:1700-1704 returns null)
        if self.isDeterministic(): return self._det(1.0 if self.valueIfNonStochastic != self.valueIfNonStochastic else 0.0)
        return self._v(capi.ISNAN)

    # ------------------------------------------------------------------ binary operators with type priority (RVC:1390-1580)
    def _add_rv(self, rv):
        if rv.getTypePriority() > self.getTypePriority(): return rv.add(self)
        newTime = max(self.time, rv.getFiltrationTime())
        if self.isDeterministic() and rv.isDeterministic(): return self._det(self.valueIfNonStochastic + rv.doubleValue(), newTime)
        if self.isDeterministic(): return self._as_cuda(rv)._vs(capi.ADD, self.valueIfNonStochastic, newTime)     # RVF:974-979
        if rv.isDeterministic(): return self._vs(capi.ADD, rv.doubleValue(), newTime)
        return self._vv(capi.ADD, self._as_cuda(rv), newTime)

    def _sub_rv(self, rv):
        if rv.getTypePriority() > self.getTypePriority(): return rv.bus(self)
        newTime = max(self.time, rv.getFiltrationTime())
        if self.isDeterministic() and rv.isDeterministic(): return self._det(self.valueIfNonStochastic - rv.doubleValue(), newTime)
        if self.isDeterministic(): return self._as_cuda(rv)._vs(capi.BUS, self.valueIfNonStochastic, newTime)     # RVF:1003-1008
        if rv.isDeterministic(): return self._vs(capi.SUB, rv.doubleValue(), newTime)
        return self._vv(capi.SUB, self._as_cuda(rv), newTime)

    def _bus_rv(self, rv):
        if rv.getTypePriority() > self.getTypePriority(): return rv.sub(self)
        newTime = max(self.time, rv.getFiltrationTime())
        if self.isDeterministic() and rv.isDeterministic(): return self._det(-self.valueIfNonStochastic + rv.doubleValue(), newTime)
        if self.isDeterministic(): return self._as_cuda(rv)._vs(capi.SUB, self.valueIfNonStochastic, newTime)     # RVF:1033-1038
        if rv.isDeterministic(): return self._vs(capi.BUS, rv.doubleValue(), newTime)
        return self._vv(capi.BUS, self._as_cuda(rv), newTime)

    def _mult_rv(self, rv):
        if rv.getTypePriority() > self.getTypePriority(): return rv.mult(self)
        newTime = max(self.time, rv.getFiltrationTime())
        if self.isDeterministic() and rv.isDeterministic(): return self._det(self.valueIfNonStochastic * rv.doubleValue(), newTime)
        if rv.isDeterministic(): return self._vs(capi.MULT, rv.doubleValue(), newTime)
        if self.isDeterministic(): return self._as_cuda(rv)._vs(capi.MULT, self.valueIfNonStochastic, newTime)    # RVF:1065-1070
        return self._vv(capi.MULT, self._as_cuda(rv), newTime)

    def _div_rv(self, rv):
        if rv.getTypePriority() > self.getTypePriority(): return rv.vid(self)
        newTime = max(self.time, rv.getFiltrationTime())
        if self.isDeterministic() and rv.isDeterministic(): return self._det(_div(self.valueIfNonStochastic, rv.doubleValue()), newTime)
        if self.isDeterministic(): return self._as_cuda(rv)._vs(capi.VID, self.valueIfNonStochastic, newTime)     # RVF:1098-1103
        if rv.isDeterministic(): return self._vs(capi.DIV, rv.doubleValue(), newTime)
        return self._vv(capi.DIV, self._as_cuda(rv), newTime)

    def _vid_rv(self, rv):
        if rv.getTypePriority() > self.getTypePriority(): return rv.div(self)       # RVF:1116-1119 (RVC:1513-1516 calls vid: defect)
        newTime = max(self.time, rv.getFiltrationTime())
        if self.isDeterministic() and rv.isDeterministic(): return self._det(_div(rv.doubleValue(), self.valueIfNonStochastic), newTime)
        if self.isDeterministic(): return self._as_cuda(rv)._vs(capi.DIV, self.valueIfNonStochastic, newTime)     # RVF:1128-1133
        if rv.isDeterministic(): return self._vs(capi.VID, rv.doubleValue(), newTime)
        return self._vv(capi.VID, self._as_cuda(rv), newTime)

    def _cap_rv(self, rv):
        if rv.getTypePriority() > self.getTypePriority(): return rv.cap(self)
        newTime = max(self.time, rv.getFiltrationTime())
        if self.isDeterministic() and rv.isDeterministic(): return self._det(_java_min(self.valueIfNonStochastic, rv.doubleValue()), newTime)
        if self.isDeterministic(): return self._as_cuda(rv)._vs(capi.CAP, self.valueIfNonStochastic, newTime)     # RVF:1158-1163
        if rv.isDeterministic(): return self._vs(capi.CAP, rv.doubleValue(), newTime)   # RVC:1546-1555 lacks this branch (NPE): defect not copied
        return self._vv(capi.CAP, self._as_cuda(rv), newTime)

    def _floor_rv(self, rv):
        if rv.getTypePriority() > self.getTypePriority(): return rv.floor(self)
        newTime = max(self.time, rv.getFiltrationTime())
        if self.isDeterministic() and rv.isDeterministic(): return self._det(_java_max(self.valueIfNonStochastic, rv.doubleValue()), newTime)
        if self.isDeterministic(): return self._as_cuda(rv)._vs(capi.FLOOR, self.valueIfNonStochastic, newTime)   # RVF:1187-1192
        if rv.isDeterministic(): return self._vs(capi.FLOOR, rv.doubleValue(), newTime)
        return self._vv(capi.FLOOR, self._as_cuda(rv), newTime)

    # ------------------------------------------------------------------ accrue / discount (RVC:1582-1624, RVF:1202-1256)
    def accrue(self, rate: RandomVariable, periodLength: float):
        if rate.getTypePriority() > self.getTypePriority(): return rate.mult(periodLength).add(1.0).mult(self)
        newTime = max(self.time, rate.getFiltrationTime())
        if rate.isDeterministic(): return self.mult(1.0 + rate.doubleValue() * periodLength)
        r = self._as_cuda(rate)
        if self.isDeterministic():                                 # RVF:1214-1219: (float)v * (1 + r*p)
            out = r._vs(capi.MULT, periodLength)._vs(capi.ADD, 1.0)._vs(capi.MULT, self.valueIfNonStochastic, newTime)
            return out
        out = C.c_uint64(0)
        capi.check(capi.load().fmc_op_vvs(capi.ACCRUE, self._h, r._h, float(periodLength), C.byref(out)))
        return self._wrap(newTime, out, self._size)

    def discount(self, rate: RandomVariable, periodLength: float):
        if rate.getTypePriority() > self.getTypePriority(): return rate.mult(periodLength).add(1.0).vid(self)   # RVF:1232-1235
        newTime = max(self.time, rate.getFiltrationTime())
        if rate.isDeterministic(): return self.div(1.0 + rate.doubleValue() * periodLength)
        r = self._as_cuda(rate)
        if self.isDeterministic():                                 # RVF:1242-1247: (float)v / (1 + r*p)
            return r._vs(capi.MULT, periodLength)._vs(capi.ADD, 1.0)._vs(capi.VID, self.valueIfNonStochastic, newTime)
        out = C.c_uint64(0)
        capi.check(capi.load().fmc_op_vvs(capi.DISCOUNT, self._h, r._h, float(periodLength), C.byref(out)))
        return self._wrap(newTime, out, self._size)

    # ------------------------------------------------------------------ ternary (RVF:1263-1438)
    def choose(self, valueIfTriggerNonNegative: RandomVariable, valueIfTriggerNegative: RandomVariable):
        a, b = valueIfTriggerNonNegative, valueIfTriggerNegative
        newTime = max(self.time, a.getFiltrationTime(), b.getFiltrationTime())
        if self.isDeterministic():                                 # RVF:1270-1276: returns the chosen object
            return a if self.valueIfNonStochastic >= 0 else b
        ca, cb = self._as_cuda(a), self._as_cuda(b)
        out = C.c_uint64(0)
        capi.check(capi.load().fmc_op_choose(
            self._h,
            ca._h, 0.0 if ca._h else ca.valueIfNonStochastic,
            cb._h, 0.0 if cb._h else cb.valueIfNonStochastic, C.byref(out)))
        return self._wrap(newTime, out, self._size)

    def addProduct(self, factor1: RandomVariable, factor2):
        if isinstance(factor2, RandomVariable):
            return self._add_product_rv(factor1, factor2)
        if factor1.getTypePriority() > self.getTypePriority(): return factor1.mult(factor2).add(self)
        newTime = max(self.time, factor1.getFiltrationTime())
        if factor1.isDeterministic(): return self.add(factor1.doubleValue() * factor2)
        f1 = self._as_cuda(factor1)
        if self.isDeterministic():                                 # RVF:1329-1334: (float)v + f1*(float)f2
            return f1._vs(capi.MULT, factor2)._vs(capi.ADD, self.valueIfNonStochastic, newTime)
        out = C.c_uint64(0)
        capi.check(capi.load().fmc_op_vvs(capi.ADDPRODUCT, self._h, f1._h, float(factor2), C.byref(out)))
        return self._wrap(newTime, out, self._size)

    def _add_product_rv(self, factor1: RandomVariable, factor2: RandomVariable):
        if factor1.getTypePriority() > self.getTypePriority() or factor2.getTypePriority() > self.getTypePriority():
            return factor1.mult(factor2).add(self)
        newTime = max(self.time, factor1.getFiltrationTime(), factor2.getFiltrationTime())
        if self.isDeterministic() and factor1.isDeterministic() and factor2.isDeterministic():
            return self._det(self.valueIfNonStochastic + factor1.doubleValue() * factor2.doubleValue(), newTime)
        if factor1.isDeterministic() and factor2.isDeterministic(): return self.add(factor1.doubleValue() * factor2.doubleValue())
        if factor2.isDeterministic(): return self.addProduct(factor1, factor2.doubleValue())
        if factor1.isDeterministic(): return self.addProduct(factor2, factor1.doubleValue())
        if not self.isDeterministic():
            f1, f2 = self._as_cuda(factor1), self._as_cuda(factor2)
            out = C.c_uint64(0)
            capi.check(capi.load().fmc_op_vvv(capi.ADDPRODUCT, self._h, f1._h, f2._h, C.byref(out)))
            return self._wrap(newTime, out, self._size)
        return self.add(factor1.mult(factor2))                     # RVF:1379-1381

    def addSumProduct(self, factor1: Sequence[RandomVariable], factor2: Sequence[RandomVariable]):   # RVF:1384-1392
        result = self
        for f1, f2 in zip(factor1, factor2):
            result = result.addProduct(f1, f2)
        return result

    def _ratio(self, numerator: RandomVariable, denominator: RandomVariable, op_vvv: int, sign: float):
        newTime = max(self.time, numerator.getFiltrationTime(), denominator.getFiltrationTime())
        if self.isDeterministic() and numerator.isDeterministic() and denominator.isDeterministic():
            return self._det(self.valueIfNonStochastic + sign * _div(numerator.doubleValue(), denominator.doubleValue()), newTime)
        n, d = self._as_cuda(numerator), self._as_cuda(denominator)
        if not self.isDeterministic() and not n.isDeterministic() and not d.isDeterministic():
            out = C.c_uint64(0)
            capi.check(capi.load().fmc_op_vvv(op_vvv, self._h, n._h, d._h, C.byref(out)))
            return self._wrap(newTime, out, self._size)
        # mixed deterministic / stochastic operands: RVF:1408-1413 evaluates (float)get(i) +- (float)n.get(i) / (float)d.get(i)
        if n.isDeterministic() and d.isDeterministic():
            with np.errstate(all="ignore"):
                ratio = float(np.float32(n.valueIfNonStochastic) / np.float32(d.valueIfNonStochastic))
            return self._vs(capi.ADD if sign > 0 else capi.SUB, ratio, newTime)
        q = n.div(d)
        res = q.add(self) if sign > 0 else q.bus(self)
        res.time = newTime
        return res

    def addRatio(self, numerator: RandomVariable, denominator: RandomVariable):
        if numerator.getTypePriority() > self.getTypePriority() or denominator.getTypePriority() > self.getTypePriority():
            return numerator.div(denominator).add(self)            # RVF:1396-1399
        return self._ratio(numerator, denominator, capi.ADDRATIO, +1.0)

    def subRatio(self, numerator: RandomVariable, denominator: RandomVariable):
        if numerator.getTypePriority() > self.getTypePriority() or denominator.getTypePriority() > self.getTypePriority():
            return numerator.div(denominator).mult(-1).add(self)   # RVF:1419-1422
        return self._ratio(numerator, denominator, capi.SUBRATIO, -1.0)

    def __repr__(self) -> str:
        if self.isDeterministic():
            return f"RandomVariableCuda(time={self.time}, value={self.valueIfNonStochastic})"
        return f"RandomVariableCuda(time={self.time}, size={self._size}, handle=0x{self._h:x})"


class RandomVariableCudaFactory:
    """RandomVariableCudaFactory (RandomVariableCudaFactory.java:18-35): the injection point of the backend."""

    def createRandomVariable(self, *args) -> RandomVariableCuda:
        # createRandomVariable(value) comes from AbstractRandomVariableFactory: time = -infinity
        if len(args) == 1:
            return RandomVariableCuda(-math.inf, args[0])
        time, value = args
        return RandomVariableCuda(time, value)                     # RVCF:26-34

    def createRandomVariableArray(self, values) -> list:
        return [self.createRandomVariable(v) for v in values]

    def __repr__(self) -> str:
        return "RandomVariableCudaFactory()"
