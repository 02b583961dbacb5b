/*
 * Conditional expectation by regression with the normal equations accumulated in ONE fused pass on the device
 * (fmc_regression_normal_eq) instead of k(k+1)/2 + k separate mult().getAverage() round trips, which is what finmath-lib's
 * MonteCarloConditionalExpectationRegression does through the RandomVariable interface (hook: RandomVariableFromFloatArray.java:861-864).
 * Optional: the library's own class works unchanged on RandomVariableCuda; this one is faster.
 *
 * Delivered as source: no JVM exists in the build environment of this repository (java/README.md).
 */
package net.finmath.cuda.montecarlo;

import org.apache.commons.math3.linear.Array2DRowRealMatrix;
import org.apache.commons.math3.linear.ArrayRealVector;
import org.apache.commons.math3.linear.SingularValueDecomposition;

import net.finmath.stochastic.ConditionalExpectationEstimator;
import net.finmath.stochastic.RandomVariable;

public class MonteCarloConditionalExpectationRegressionCuda implements ConditionalExpectationEstimator {

	private final RandomVariable[] basisFunctionsEstimator;
	private final RandomVariable[] basisFunctionsPredictor;

	public MonteCarloConditionalExpectationRegressionCuda(final RandomVariable[] basisFunctions) {
		this(basisFunctions, basisFunctions);
	}

	public MonteCarloConditionalExpectationRegressionCuda(final RandomVariable[] basisFunctionsEstimator, final RandomVariable[] basisFunctionsPredictor) {
		this.basisFunctionsEstimator = basisFunctionsEstimator;
		this.basisFunctionsPredictor = basisFunctionsPredictor;
	}

	@Override
	public RandomVariable getConditionalExpectation(final RandomVariable randomVariable) {
		final double[] linearRegressionParameters = getLinearRegressionParameters(randomVariable);
		RandomVariable conditionalExpectation = basisFunctionsPredictor[0].mult(linearRegressionParameters[0]);
		for(int i = 1; i < basisFunctionsPredictor.length; i++) {
			conditionalExpectation = conditionalExpectation.addProduct(basisFunctionsPredictor[i], linearRegressionParameters[i]);
		}
		return conditionalExpectation;
	}

	/** Solves X^T X c = X^T y, both sides accumulated in one pass over the basis functions and y (float products, double sums). */
	public double[] getLinearRegressionParameters(final RandomVariable dependents) {
		final int k = basisFunctionsEstimator.length;
		final long[] handles = new long[k];
		final double[] scalars = new double[k];
		final RandomVariableCuda[] keepAlive = new RandomVariableCuda[k];
		for(int i = 0; i < k; i++) {
			if(basisFunctionsEstimator[i].isDeterministic()) {
				handles[i] = 0L;
				scalars[i] = basisFunctionsEstimator[i].doubleValue();
			}
			else {
				keepAlive[i] = basisFunctionsEstimator[i] instanceof RandomVariableCuda ? (RandomVariableCuda) basisFunctionsEstimator[i]
						: new RandomVariableCuda(basisFunctionsEstimator[i]);
				handles[i] = keepAlive[i].getHandle();
			}
		}
		final RandomVariableCuda y = dependents instanceof RandomVariableCuda ? (RandomVariableCuda) dependents : new RandomVariableCuda(dependents);
		final double[] xtx = new double[k * k];
		final double[] xty = new double[k];
		FmCuda.regressionNormalEquations(handles, scalars, y.getHandle(), xtx, xty);
		final double[][] matrix = new double[k][k];
		for(int i = 0; i < k; i++) {
			System.arraycopy(xtx, i * k, matrix[i], 0, k);
		}
		return new SingularValueDecomposition(new Array2DRowRealMatrix(matrix, false)).getSolver().solve(new ArrayRealVector(xty, false)).toArray();
	}
}
