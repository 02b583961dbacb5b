"""MonteCarloConditionalExpectationRegression on the device.

finmath-lib's estimator (net.finmath.montecarlo.conditionalexpectation.MonteCarloConditionalExpectationRegression,
not vendored in the reference; reached through RandomVariable.getConditionalExpectation(estimator),
RandomVariableFromFloatArray.java:860-864) solves the normal equations
    (X^T X) c = X^T y,   XtX[i][j] = E[b_i b_j],  XtY[i] = E[y b_i]
and returns sum_i c_i * b_i. The reference path needs k(k+1)/2 + k separate mult + getAverage round trips, each a
full device->host copy (RandomVariableCuda.java:869-883); here one fused kernel produces all sums
(fmc_regression_normal_eq), the k x k solve (k <= 12) is done on the host in double like commons-math3 does.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _capi as capi
from .random_variable import RandomVariable, RandomVariableCuda


def normal_equations(basis: Sequence[RandomVariable], y: RandomVariable):
    """Returns (XtX [k,k], XtY [k]) as float64 arrays."""
    capi.ensure_init()
    k = len(basis)
    cb = [RandomVariableCuda._as_cuda(b) for b in basis]
    cy = RandomVariableCuda._as_cuda(y)
    if cy.isDeterministic():
        raise ValueError("the dependent variable must be stochastic")
    handles = (C.c_uint64 * k)(*[b.handle for b in cb])
    scalars = (C.c_double * k)(*[0.0 if b.handle else b.valueIfNonStochastic for b in cb])
    XtX = np.empty((k, k), dtype=np.float64)
    XtY = np.empty(k, dtype=np.float64)
    capi.check(capi.load().fmc_regression_normal_eq(handles, scalars, k, cy.handle, XtX.ctypes.data, XtY.ctypes.data))
    return XtX, XtY


class MonteCarloConditionalExpectationRegression:
    def __init__(self, basisFunctionsEstimator: Sequence[RandomVariable], basisFunctionsPredictor: Sequence[RandomVariable] | None = None):
        self.basisFunctionsEstimator = list(basisFunctionsEstimator)
        self.basisFunctionsPredictor = list(basisFunctionsPredictor) if basisFunctionsPredictor is not None else self.basisFunctionsEstimator

    def getLinearRegressionParameters(self, dependents: RandomVariable) -> np.ndarray:
        XtX, XtY = normal_equations(self.basisFunctionsEstimator, dependents)
        # commons-math3 SingularValueDecomposition(XTX).getSolver().solve(XTY): minimum-norm least squares
        return np.linalg.lstsq(XtX, XtY, rcond=1e-10)[0]

    def getConditionalExpectation(self, randomVariable: RandomVariable) -> RandomVariable:
        coeff = self.getLinearRegressionParameters(randomVariable)
        basis = self.basisFunctionsPredictor
        cond = basis[0].mult(float(coeff[0]))
        for i in range(1, len(basis)):
            cond = cond.addProduct(basis[i], float(coeff[i]))
        return cond
