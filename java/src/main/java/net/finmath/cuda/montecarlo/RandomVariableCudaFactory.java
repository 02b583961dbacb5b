/*
 * RandomVariableFactory creating RandomVariableCuda objects: the injection point of the backend, unchanged from the reference
 * (finmath-lib-cuda-extensions, RandomVariableCudaFactory.java:18-35).
 */
package net.finmath.cuda.montecarlo;

import net.finmath.montecarlo.AbstractRandomVariableFactory;
import net.finmath.montecarlo.RandomVariableFactory;
import net.finmath.stochastic.RandomVariable;

/**
 * RandomVariableFactory creating CUDA random variables (objects implementing RandomVariable whose realizations live on the device).
 */
public class RandomVariableCudaFactory extends AbstractRandomVariableFactory implements RandomVariableFactory {

	private static final long serialVersionUID = 1L;

	public RandomVariableCudaFactory() {
		super();
	}

	@Override
	public RandomVariable createRandomVariable(final double time, final double value) {
		return new RandomVariableCuda(time, value);
	}

	@Override
	public RandomVariable createRandomVariable(final double time, final double[] values) {
		return new RandomVariableCuda(time, values);
	}
}
