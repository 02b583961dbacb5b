"""finmath_cuda — B200-native backend for finmath-lib's RandomVariable / BrownianMotion vector layer.

Host-side mirror (Python) of the reference's plug-in surface over the C ABI of include/fmcuda.h:
RandomVariableCudaFactory, RandomVariableCuda (type priority 20), BrownianMotionCuda,
MonteCarloConditionalExpectationRegression. No CPU fallback: importing works without a GPU (so that the symbol
table can be checked), creating a stochastic vector without one raises.
"""
from . import _capi
from ._capi import CudaError, ensure_init, flush, pool_trim, reset_stats, set_option, shutdown, stats, sync
from .random_variable import RandomVariable, RandomVariableCuda, RandomVariableCudaFactory
from .brownian_motion import BrownianMotionCuda, TimeDiscretization
from .conditional_expectation import MonteCarloConditionalExpectationRegression
from .differentiable import RandomVariableDifferentiableAAD, RandomVariableDifferentiableAADFactory
from . import distributed

__all__ = [
    "RandomVariable", "RandomVariableCuda", "RandomVariableCudaFactory", "BrownianMotionCuda", "TimeDiscretization",
    "MonteCarloConditionalExpectationRegression", "RandomVariableDifferentiableAAD", "RandomVariableDifferentiableAADFactory", "CudaError", "ensure_init", "shutdown", "stats", "set_option", "sync", "flush", "pool_trim", "reset_stats", "distributed",
]
