/*
 * Brownian motion whose increments are generated ON the device with the random number stream of finmath-lib's
 * BrownianMotionFromMersenneRandomNumbers (commons-math3 MersenneTwister, nextDouble() = 26 + 26 bits, inverse normal AS241):
 * the same numbers the CPU class produces for the same seed, so that a model run with RandomVariableCudaFactory reproduces the
 * CPU run path by path. Replaces the cuRAND based BrownianMotionCudaWithRandomVariableCuda of the reference
 * (alternative/BrownianMotionCudaWithRandomVariableCuda.java:49-255, whose XORWOW stream has no CPU counterpart).
 *
 * Delivered as source: no JVM exists in the build environment of this repository (java/README.md).
 */
package net.finmath.cuda.montecarlo;

import java.io.Serializable;
import java.util.Objects;

import net.finmath.montecarlo.BrownianMotion;
import net.finmath.montecarlo.RandomVariableFactory;
import net.finmath.stochastic.RandomVariable;
import net.finmath.time.TimeDiscretization;

/**
 * Implementation of a time-discrete n-dimensional Brownian motion W = (W_1, ..., W_n) with independent components, generated on the
 * CUDA device. The increments are created lazily on first access, are kept in device memory and are never copied to the host.
 */
public class BrownianMotionCuda implements BrownianMotion, Serializable {

	private static final long serialVersionUID = -5430067621669213475L;

	/** How the Mersenne Twister is seeded: as net.finmath.randomnumbers.MersenneTwister(long) does (finmath-lib 5.x), or as commons-math3 MersenneTwister(int). */
	public enum SeedMode { LONG, INT }

	private final TimeDiscretization	timeDiscretization;
	private final int					numberOfFactors;
	private final int					numberOfPaths;
	private final int					seed;
	private final SeedMode				seedMode;

	private final RandomVariableFactory	randomVariableFactory;		// for getRandomVariableForConstant only (the increments live on the device)

	private transient RandomVariable[][]	brownianIncrements;			// [timeIndex][factor]; transient like BrownianMotionCudaWithRandomVariableCuda.java:61
	private final Object					brownianIncrementsLazyInitLock = new Object[0];	// a serializable lock object

	public BrownianMotionCuda(final TimeDiscretization timeDiscretization, final int numberOfFactors, final int numberOfPaths, final int seed,
			final RandomVariableFactory randomVariableFactory, final SeedMode seedMode) {
		this.timeDiscretization = timeDiscretization;
		this.numberOfFactors = numberOfFactors;
		this.numberOfPaths = numberOfPaths;
		this.seed = seed;
		this.seedMode = seedMode;
		this.randomVariableFactory = randomVariableFactory;
		this.brownianIncrements = null;	// lazy initialization
	}

	public BrownianMotionCuda(final TimeDiscretization timeDiscretization, final int numberOfFactors, final int numberOfPaths, final int seed,
			final RandomVariableFactory randomVariableFactory) {
		this(timeDiscretization, numberOfFactors, numberOfPaths, seed, randomVariableFactory, SeedMode.LONG);
	}

	public BrownianMotionCuda(final TimeDiscretization timeDiscretization, final int numberOfFactors, final int numberOfPaths, final int seed) {
		this(timeDiscretization, numberOfFactors, numberOfPaths, seed, new RandomVariableCudaFactory());
	}

	@Override
	public BrownianMotion getCloneWithModifiedSeed(final int seed) {
		return new BrownianMotionCuda(getTimeDiscretization(), getNumberOfFactors(), getNumberOfPaths(), seed, randomVariableFactory, seedMode);
	}

	@Override
	public BrownianMotion getCloneWithModifiedTimeDiscretization(final TimeDiscretization newTimeDiscretization) {
		return new BrownianMotionCuda(newTimeDiscretization, getNumberOfFactors(), getNumberOfPaths(), getSeed(), randomVariableFactory, seedMode);
	}

	@Override
	public RandomVariable getBrownianIncrement(final int timeIndex, final int factor) {
		// Thread safe lazy initialization
		synchronized(brownianIncrementsLazyInitLock) {
			if(brownianIncrements == null) {
				doGenerateBrownianMotion();
			}
		}
		return brownianIncrements[timeIndex][factor];
	}

	/**
	 * One call into the runtime generates all increments: every GPU block jumps the Mersenne Twister ahead to its part of the
	 * path-major stream (path outer, then time index, then factor: the loop order of BrownianMotionFromMersenneRandomNumbers), draws
	 * two 32 bit words per uniform, applies the inverse normal in double, scales by sqrt(dt) and rounds to float.
	 */
	private void doGenerateBrownianMotion() {
		final int numberOfTimeSteps = timeDiscretization.getNumberOfTimeSteps();
		final double[] sqrtOfTimeStep = new double[numberOfTimeSteps];
		for(int timeIndex = 0; timeIndex < numberOfTimeSteps; timeIndex++) {
			sqrtOfTimeStep[timeIndex] = Math.sqrt(timeDiscretization.getTimeStep(timeIndex));
		}
		final long[] handles = FmCuda.brownianIncrements(seedMode == SeedMode.LONG ? 0 : 1, seed, numberOfTimeSteps, numberOfFactors, 0L, numberOfPaths, sqrtOfTimeStep);
		final RandomVariable[][] increments = new RandomVariable[numberOfTimeSteps][numberOfFactors];
		for(int timeIndex = 0; timeIndex < numberOfTimeSteps; timeIndex++) {
			final double time = timeDiscretization.getTime(timeIndex + 1);
			for(int factor = 0; factor < numberOfFactors; factor++) {
				increments[timeIndex][factor] = RandomVariableCuda.of(time, handles[timeIndex * numberOfFactors + factor], numberOfPaths);
			}
		}
		brownianIncrements = increments;
	}

	@Override
	public TimeDiscretization getTimeDiscretization() {
		return timeDiscretization;
	}

	@Override
	public int getNumberOfFactors() {
		return numberOfFactors;
	}

	@Override
	public int getNumberOfPaths() {
		return numberOfPaths;
	}

	@Override
	public RandomVariable getRandomVariableForConstant(final double value) {
		return randomVariableFactory.createRandomVariable(value);
	}

	public int getSeed() {
		return seed;
	}

	@Override
	public RandomVariable getIncrement(final int timeIndex, final int factor) {
		return getBrownianIncrement(timeIndex, factor);
	}

	@Override
	public String toString() {
		return super.toString() + "\n" + "timeDiscretization: " + timeDiscretization.toString() + "\n" + "numberOfPaths: " + numberOfPaths + "\n"
				+ "numberOfFactors: " + numberOfFactors + "\n" + "seed: " + seed + " (" + seedMode + ")";
	}

	@Override
	public boolean equals(final Object o) {
		if(this == o) {
			return true;
		}
		if(o == null || getClass() != o.getClass()) {
			return false;
		}
		final BrownianMotionCuda that = (BrownianMotionCuda) o;
		return numberOfFactors == that.numberOfFactors && numberOfPaths == that.numberOfPaths && seed == that.seed && seedMode == that.seedMode
				&& timeDiscretization.equals(that.timeDiscretization);
	}

	@Override
	public int hashCode() {
		return Objects.hash(timeDiscretization, numberOfFactors, numberOfPaths, seed, seedMode);
	}
}
