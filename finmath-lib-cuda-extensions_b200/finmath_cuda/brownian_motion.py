"""BrownianMotionCuda — Brownian increments generated on the device.

API shape: net.finmath.montecarlo.BrownianMotion as implemented by the reference's
BrownianMotionCudaWithRandomVariableCuda (/root/reference/src/main/java/net/finmath/cuda/montecarlo/alternative/
BrownianMotionCudaWithRandomVariableCuda.java:79-250): lazy initialisation under a lock, one RandomVariable per
(timeIndex, factor) with filtration time t_{i+1}, clone-with-seed, getRandomVariableForConstant.
Random stream: NOT cuRAND XORWOW (BMC:159) but the stream of finmath-lib's BrownianMotionFromMersenneRandomNumbers
(commons-math3 MT19937, nextDouble, AS241 inverse normal, path-major draw order), which is what the reference's
heavy tests feed to RandomVariableCudaFactory (LIBORMarketModelCalibrationATMTest.java:283).
"""
from __future__ import annotations

import ctypes as C
import math
import threading

import numpy as np

from . import _capi as capi
from .random_variable import RandomVariable, RandomVariableCuda

SEED_MODE_LONG = 0     # net.finmath.randomnumbers.MersenneTwister(long) -> commons-math3 setSeed(long) -> init_by_array{hi, lo}
SEED_MODE_INT = 1      # commons-math3 MersenneTwister(int) -> init_genrand


class TimeDiscretization:
    """Minimal net.finmath.time.TimeDiscretizationFromArray: (initial, numberOfTimeSteps, deltaT) or an array of times."""

    def __init__(self, initial_or_times, numberOfTimeSteps: int | None = None, deltaT: float | None = None):
        if numberOfTimeSteps is None:
            self.times = np.asarray(initial_or_times, dtype=np.float64).copy()
        else:
            self.times = float(initial_or_times) + np.arange(numberOfTimeSteps + 1, dtype=np.float64) * float(deltaT)

    def getNumberOfTimeSteps(self) -> int: return self.times.size - 1
    def getNumberOfTimes(self) -> int: return self.times.size
    def getTime(self, i: int) -> float: return float(self.times[i])
    def getTimeStep(self, i: int) -> float: return float(self.times[i + 1] - self.times[i])

    def getTimeIndex(self, time: float) -> int:
        i = int(np.searchsorted(self.times, time))
        if i < self.times.size and abs(self.times[i] - time) < 1e-12:
            return i
        if i > 0 and abs(self.times[i - 1] - time) < 1e-12:
            return i - 1
        return -i - 1                                              # java.util.Arrays.binarySearch convention

    def __eq__(self, other): return isinstance(other, TimeDiscretization) and np.array_equal(self.times, other.times)
    def __hash__(self): return hash(self.times.tobytes())


class BrownianMotionCuda:
    def __init__(self, timeDiscretization: TimeDiscretization, numberOfFactors: int, numberOfPaths: int, seed: int,
                 randomVariableFactory=None, seedMode: int = SEED_MODE_LONG, pathRange: tuple[int, int] | None = None):
        """pathRange = (p0, p1): this process holds only paths [p0, p1) of the numberOfPaths-path motion (multi-GPU
        sharding; the increments are those the single-process motion has on the same paths)."""
        self.timeDiscretization = timeDiscretization
        self.numberOfFactors = int(numberOfFactors)
        self.numberOfPaths = int(numberOfPaths)
        self.seed = int(seed)
        self.seedMode = int(seedMode)
        self.pathRange = (0, self.numberOfPaths) if pathRange is None else (int(pathRange[0]), int(pathRange[1]))
        self.randomVariableFactory = randomVariableFactory         # kept for signature compatibility (BMC:90 ignores it too)
        self._increments = None                                    # transient, lazily generated (BMC:61, 123-136)
        self._lock = threading.Lock()

    def getCloneWithModifiedSeed(self, seed: int) -> "BrownianMotionCuda":                      # BMC:111-114
        return BrownianMotionCuda(self.timeDiscretization, self.numberOfFactors, self.numberOfPaths, seed,
                                  self.randomVariableFactory, self.seedMode, self.pathRange)

    def getCloneWithModifiedTimeDiscretization(self, newTimeDiscretization) -> "BrownianMotionCuda":   # BMC:116-120
        return BrownianMotionCuda(newTimeDiscretization, self.numberOfFactors, self.numberOfPaths, self.seed,
                                  self.randomVariableFactory, self.seedMode, self.pathRange)

    def getBrownianIncrement(self, timeIndex: int, factor: int) -> RandomVariable:              # BMC:122-136
        with self._lock:
            if self._increments is None:
                self._generate()
        return self._increments[timeIndex][factor]

    getIncrement = getBrownianIncrement                                                          # BMC:247-250

    def _generate(self) -> None:                                                                 # BMC:141-182
        capi.ensure_init()
        td = self.timeDiscretization
        T, F = td.getNumberOfTimeSteps(), self.numberOfFactors
        sqrt_dt = np.array([math.sqrt(td.getTimeStep(t)) for t in range(T)], dtype=np.float64)
        handles = (C.c_uint64 * (T * F))()
        p0, p1 = self.pathRange
        capi.check(capi.load().fmc_brownian_generate(self.seedMode, self.seed, T, F, p0, p1, sqrt_dt.ctypes.data, handles))
        n = p1 - p0
        self._increments = [[RandomVariableCuda.of(td.getTime(t + 1), handles[t * F + f], n) for f in range(F)] for t in range(T)]

    def getTimeDiscretization(self) -> TimeDiscretization: return self.timeDiscretization       # BMC:184-187
    def getNumberOfFactors(self) -> int: return self.numberOfFactors                             # BMC:189-192
    def getNumberOfPaths(self) -> int: return self.numberOfPaths                                 # BMC:194-197
    def getRandomVariableForConstant(self, value: float) -> RandomVariable:                      # BMC:199-202
        return RandomVariableCuda(value)
    def getSeed(self) -> int: return self.seed                                                   # BMC:207-209

    def __eq__(self, o):                                                                         # BMC:220-245
        return (isinstance(o, BrownianMotionCuda) and self.numberOfFactors == o.numberOfFactors and self.numberOfPaths == o.numberOfPaths
                and self.seed == o.seed and self.timeDiscretization == o.timeDiscretization)

    def __hash__(self):                                                                          # BMC:252-259
        r = hash(self.timeDiscretization)
        for v in (self.numberOfFactors, self.numberOfPaths, self.seed):
            r = (31 * r + v) & 0xffffffff
        return r


def mt19937_raw(seed: int, count: int, seedMode: int = SEED_MODE_LONG, skip: int = 0) -> np.ndarray:
    """Tempered uint32 words [skip, skip+count) of the device generator (bit-exactness witness)."""
    capi.ensure_init()
    out = np.empty(int(count), dtype=np.uint32)
    capi.check(capi.load().fmc_mt19937_raw(int(seedMode), int(seed), int(skip), int(count), out.ctypes.data))
    return out
