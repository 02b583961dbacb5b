/*
 * RandomVariableCuda over the B200-native runtime (include/fmcuda.h). Same public surface as the reference class
 * (finmath-lib-cuda-extensions, RandomVariableCuda.java:566-577, 618-757, 785-1701): immutable, thread safe, type priority 20,
 * deterministic values kept on the host as doubles. Behaviour follows the CPU twin RandomVariableFromFloatArray ("RVF") wherever
 * the reference GPU class is incomplete or defective (choose, isNaN, sin, cos, cap with a deterministic argument, getQuantile,
 * getVariance; see SURVEY.md Appendix B of this repository).
 *
 * Delivered as source: no JVM exists in the build environment of this repository (java/README.md).
 */
package net.finmath.cuda.montecarlo;

import java.io.IOException;
import java.io.ObjectInputStream;
import java.io.ObjectOutputStream;
import java.lang.ref.Cleaner;
import java.util.Arrays;
import java.util.function.DoubleBinaryOperator;
import java.util.function.DoubleUnaryOperator;
import java.util.function.IntToDoubleFunction;
import java.util.stream.DoubleStream;

import net.finmath.functions.DoubleTernaryOperator;
import net.finmath.stochastic.RandomVariable;

/**
 * A random variable whose realizations live in the memory of a CUDA device as 32 bit floats. Operations are recorded by the native
 * runtime and executed, fused, when a value is demanded (a reduction, getRealizations, get): every operand crosses device memory once.
 */
public class RandomVariableCuda implements RandomVariable {

	private static final long serialVersionUID = 7620120320663270600L;		// RandomVariableCuda.java:564

	private static final int typePriorityDefault = 20;						// RandomVariableCuda.java:568

	private final double	time;					// filtration time
	private final int		typePriority;
	private final long		size;
	private final double	valueIfNonStochastic;	// the value if handle == 0

	// the device vector (0: deterministic). Not serialized: writeObject stores the realizations, readObject uploads them again.
	private transient long					handle;
	private transient Cleaner.Cleanable		cleanable;

	/* ------------------------------------------------------------------ construction (RandomVariableCuda.java:618-734) */

	private RandomVariableCuda(final double time, final long handle, final long size, final int typePriority) {
		this.time = time;
		this.typePriority = typePriority;
		this.size = size;
		this.valueIfNonStochastic = Double.NaN;
		attach(handle);
	}

	private void attach(final long newHandle) {
		handle = newHandle;
		// the cleaning action must not capture this: only the handle
		cleanable = newHandle != 0L ? FmCuda.CLEANER.register(this, () -> FmCuda.release(newHandle)) : null;
	}

	/** Wrap a device vector produced by the runtime (the caller's reference to the handle passes to the new object). */
	static RandomVariableCuda of(final double time, final long handle, final long size, final int typePriority) {
		return new RandomVariableCuda(time, handle, size, typePriority);
	}

	static RandomVariableCuda of(final double time, final long handle, final long size) {
		return new RandomVariableCuda(time, handle, size, typePriorityDefault);
	}

	public static RandomVariableCuda of(final double time, final double value) {
		return new RandomVariableCuda(time, value);
	}

	public RandomVariableCuda(final double value) {
		this(-Double.MAX_VALUE, value);
	}

	public RandomVariableCuda(final double time, final double value) {
		this(time, value, typePriorityDefault);
	}

	public RandomVariableCuda(final double time, final double value, final int typePriority) {
		this.time = time;
		this.typePriority = typePriority;
		this.size = 1;
		this.valueIfNonStochastic = value;
		this.handle = 0L;
	}

	public RandomVariableCuda(final double time, final float[] realisations, final int typePriority) {
		this.time = time;
		this.typePriority = typePriority;
		this.size = realisations.length;
		this.valueIfNonStochastic = Double.NaN;
		attach(FmCuda.upload(realisations));
	}

	public RandomVariableCuda(final double time, final float[] realisations) {
		this(time, realisations, typePriorityDefault);
	}

	public RandomVariableCuda(final double time, final double[] realisations) {
		this.time = time;
		this.typePriority = typePriorityDefault;
		this.size = realisations.length;
		this.valueIfNonStochastic = Double.NaN;
		attach(FmCuda.upload(realisations));									// (float) cast in the runtime, RandomVariableCuda.java:768-774
	}

	public RandomVariableCuda(final float[] realisations) {
		this(0.0, realisations);
	}

	/** Copy of any other implementation of RandomVariable (RVF:64-70). */
	public RandomVariableCuda(final RandomVariable value) {
		this.time = value.getFiltrationTime();
		this.typePriority = typePriorityDefault;
		if(value.isDeterministic()) {
			this.size = 1;
			this.valueIfNonStochastic = value.doubleValue();
			this.handle = 0L;
		}
		else {
			final double[] realizations = value.getRealizations();
			this.size = realizations.length;
			this.valueIfNonStochastic = Double.NaN;
			attach(FmCuda.upload(realizations));
		}
	}

	/** Return cached device memory to the driver (RandomVariableCuda.java:751-753). */
	public static void clean() {
		FmCuda.poolTrim();
	}

	/** Release everything the pool holds (RandomVariableCuda.java:755-757). */
	public static void purge() {
		FmCuda.poolPurge();
	}

	/* ------------------------------------------------------------------ serialization (RandomVariableCuda.java:564: the class is Serializable) */

	private void writeObject(final ObjectOutputStream out) throws IOException {
		out.defaultWriteObject();
		out.writeObject(isDeterministic() ? null : FmCuda.downloadAsFloat(handle, size));
	}

	private void readObject(final ObjectInputStream in) throws IOException, ClassNotFoundException {
		in.defaultReadObject();
		final float[] realizations = (float[]) in.readObject();
		attach(realizations != null ? FmCuda.upload(realizations) : 0L);
	}

	/* ------------------------------------------------------------------ helpers */

	long getHandle() {
		return handle;
	}

	/** The device vector of a stochastic operand of any type; a foreign (CPU) random variable is uploaded (RandomVariableCuda.java:759-766). */
	private static RandomVariableCuda onDevice(final RandomVariable randomVariable) {
		if(randomVariable instanceof RandomVariableCuda) {
			return (RandomVariableCuda) randomVariable;
		}
		return new RandomVariableCuda(randomVariable.getFiltrationTime(), randomVariable.getRealizations());
	}

	private RandomVariableCuda result(final double newTime, final long newHandle) {
		return of(newTime, newHandle, size, typePriority);
	}

	private RandomVariableCuda scalarOp(final int opcode, final double value) {
		return result(time, FmCuda.opVS(opcode, handle, value));
	}

	private RandomVariableCuda unaryOp(final int opcode) {
		return result(time, FmCuda.opV(opcode, handle));
	}

	/* ------------------------------------------------------------------ value access */

	@Override
	public boolean equals(final RandomVariable randomVariable) {
		if(this.time != randomVariable.getFiltrationTime()) {
			return false;
		}
		if(this.isDeterministic() && randomVariable.isDeterministic()) {
			return this.valueIfNonStochastic == randomVariable.doubleValue();
		}
		if(this.isDeterministic() != randomVariable.isDeterministic()) {
			return false;
		}
		return Arrays.equals(getRealizations(), randomVariable.getRealizations());			// RVF:240-263 (the reference GPU class throws)
	}

	@Override
	public double getFiltrationTime() {
		return time;
	}

	@Override
	public int getTypePriority() {
		return typePriority;
	}

	@Override
	public double get(final int pathOrState) {
		return isDeterministic() ? valueIfNonStochastic : FmCuda.get(handle, pathOrState);		// RVF:266-272
	}

	@Override
	public int size() {
		return (int) size;
	}

	@Override
	public boolean isDeterministic() {
		return handle == 0L;
	}

	@Override
	public RandomVariable cache() {
		return this;
	}

	@Override
	public double[] getRealizations() {
		if(isDeterministic()) {
			return new double[] { valueIfNonStochastic };
		}
		return FmCuda.downloadAsDouble(handle, size);
	}

	@Override
	public Double doubleValue() {
		if(isDeterministic()) {
			return valueIfNonStochastic;
		}
		throw new UnsupportedOperationException("The random variable is non-deterministic");
	}

	@Override
	public IntToDoubleFunction getOperator() {
		if(isDeterministic()) {
			return i -> valueIfNonStochastic;
		}
		final double[] realizations = getRealizations();
		return i -> realizations[i];
	}

	@Override
	public DoubleStream getRealizationsStream() {
		return isDeterministic() ? DoubleStream.generate(() -> valueIfNonStochastic) : Arrays.stream(getRealizations());
	}

	/* ------------------------------------------------------------------ statistics (RandomVariableCuda.java:830-1091; definitions of RVF:284-602) */

	@Override
	public double getMin() {
		return isDeterministic() ? valueIfNonStochastic : FmCuda.reduce(FmCuda.RED_MIN, handle, 0L);
	}

	@Override
	public double getMax() {
		return isDeterministic() ? valueIfNonStochastic : FmCuda.reduce(FmCuda.RED_MAX, handle, 0L);
	}

	@Override
	public double getAverage() {
		return isDeterministic() ? valueIfNonStochastic : FmCuda.reduce(FmCuda.RED_AVERAGE, handle, 0L);
	}

	@Override
	public double getAverage(final RandomVariable probabilities) {
		if(isDeterministic()) {
			return valueIfNonStochastic * probabilities.getAverage();
		}
		if(probabilities.isDeterministic()) {
			return mult(probabilities.doubleValue()).getAverage();
		}
		final RandomVariableCuda weights = onDevice(probabilities);
		return FmCuda.reduce(FmCuda.RED_AVERAGE_W, handle, weights.handle);
	}

	@Override
	public double getVariance() {
		return isDeterministic() ? 0.0 : FmCuda.reduce(FmCuda.RED_VARIANCE, handle, 0L);
	}

	@Override
	public double getVariance(final RandomVariable probabilities) {
		if(isDeterministic()) {
			return 0.0;
		}
		final RandomVariableCuda weights = onDevice(probabilities.isDeterministic() ? sub(this).add(probabilities.doubleValue()) : probabilities);
		return FmCuda.reduce(FmCuda.RED_VARIANCE_W, handle, weights.handle);
	}

	@Override
	public double getSampleVariance() {
		return isDeterministic() ? 0.0 : FmCuda.reduce(FmCuda.RED_SAMPLE_VARIANCE, handle, 0L);
	}

	@Override
	public double getStandardDeviation() {
		return isDeterministic() ? 0.0 : Math.sqrt(getVariance());
	}

	@Override
	public double getStandardDeviation(final RandomVariable probabilities) {
		return isDeterministic() ? 0.0 : Math.sqrt(getVariance(probabilities));
	}

	@Override
	public double getStandardError() {
		return isDeterministic() ? 0.0 : getStandardDeviation() / Math.sqrt(size());
	}

	@Override
	public double getStandardError(final RandomVariable probabilities) {
		return isDeterministic() ? 0.0 : getStandardDeviation(probabilities) / Math.sqrt(size());
	}

	@Override
	public double getQuantile(final double quantile) {
		return isDeterministic() ? valueIfNonStochastic : FmCuda.quantile(handle, quantile);		// RVF:473-487 (not the 1-quantile of RandomVariableCuda.java:983)
	}

	@Override
	public double getQuantile(final double quantile, final RandomVariable probabilities) {
		throw new UnsupportedOperationException("Not implemented.");								// as RVF:489-499 and RandomVariableCuda.java:989-998
	}

	@Override
	public double getQuantileExpectation(final double quantileStart, final double quantileEnd) {
		if(isDeterministic()) {
			return valueIfNonStochastic;
		}
		if(quantileStart > quantileEnd) {
			return getQuantileExpectation(quantileEnd, quantileStart);
		}
		return FmCuda.quantileExpectation(handle, quantileStart, quantileEnd);
	}

	@Override
	public double[] getHistogram(final double[] intervalPoints) {
		if(isDeterministic()) {
			final double[] histogramValues = new double[intervalPoints.length + 1];
			java.util.Arrays.fill(histogramValues, 0.0);
			for(int intervalIndex = 0; intervalIndex < intervalPoints.length; intervalIndex++) {
				if(valueIfNonStochastic > intervalPoints[intervalIndex]) {
					histogramValues[intervalIndex] = 1.0;
					return histogramValues;
				}
			}
			histogramValues[intervalPoints.length] = 1.0;
			return histogramValues;
		}
		return FmCuda.histogram(handle, intervalPoints);
	}

	@Override
	public double[][] getHistogram(final int numberOfPoints, final double standardDeviations) {
		final double[] intervalPoints = new double[numberOfPoints];
		final double[] anchorPoints = new double[numberOfPoints + 1];
		final double center = getAverage();
		final double radius = standardDeviations * getStandardDeviation();
		final double stepSize = (numberOfPoints - 1) / 2.0;
		for(int i = 0; i < numberOfPoints; i++) {
			final double alpha = (-(double) (numberOfPoints - 1) / 2.0 + i) / stepSize;
			intervalPoints[i] = center + alpha * radius;
			anchorPoints[i] = center + alpha * radius - radius / (2 * stepSize);
		}
		anchorPoints[numberOfPoints] = center + 1 * radius + radius / (2 * stepSize);
		final double[][] result = new double[2][];
		result[0] = anchorPoints;
		result[1] = getHistogram(intervalPoints);
		return result;
	}

	/* ------------------------------------------------------------------ generic operators: evaluated on the host (no device counterpart of a Java lambda) */

	@Override
	public RandomVariable apply(final DoubleUnaryOperator function) {
		if(isDeterministic()) {
			return of(time, function.applyAsDouble(valueIfNonStochastic));
		}
		final double[] values = getRealizations();
		for(int i = 0; i < values.length; i++) {
			values[i] = function.applyAsDouble(values[i]);
		}
		return new RandomVariableCuda(time, values);
	}

	@Override
	public RandomVariable apply(final DoubleBinaryOperator operator, final RandomVariable argument) {
		final double newTime = Math.max(time, argument.getFiltrationTime());
		if(isDeterministic() && argument.isDeterministic()) {
			return of(newTime, operator.applyAsDouble(valueIfNonStochastic, argument.doubleValue()));
		}
		final int n = Math.max(size(), argument.size());
		final double[] values = new double[n];
		final IntToDoubleFunction a = getOperator(), b = argument.getOperator();
		for(int i = 0; i < n; i++) {
			values[i] = operator.applyAsDouble(a.applyAsDouble(i), b.applyAsDouble(i));
		}
		return new RandomVariableCuda(newTime, values);
	}

	@Override
	public RandomVariable apply(final DoubleTernaryOperator operator, final RandomVariable argument1, final RandomVariable argument2) {
		final double newTime = Math.max(time, Math.max(argument1.getFiltrationTime(), argument2.getFiltrationTime()));
		final int n = Math.max(size(), Math.max(argument1.size(), argument2.size()));
		if(n == 1) {
			return of(newTime, operator.applyAsDouble(get(0), argument1.get(0), argument2.get(0)));
		}
		final double[] values = new double[n];
		final IntToDoubleFunction a = getOperator(), b = argument1.getOperator(), c = argument2.getOperator();
		for(int i = 0; i < n; i++) {
			values[i] = operator.applyAsDouble(a.applyAsDouble(i), b.applyAsDouble(i), c.applyAsDouble(i));
		}
		return new RandomVariableCuda(newTime, values);
	}

	/* ------------------------------------------------------------------ operations with a scalar (RandomVariableCuda.java:1172-1277; RVF:751-850) */

	@Override
	public RandomVariable cap(final double cap) {
		return isDeterministic() ? of(time, Math.min(valueIfNonStochastic, cap)) : scalarOp(FmCuda.CAP, cap);
	}

	@Override
	public RandomVariable floor(final double floor) {
		return isDeterministic() ? of(time, Math.max(valueIfNonStochastic, floor)) : scalarOp(FmCuda.FLOOR, floor);
	}

	@Override
	public RandomVariable add(final double value) {
		return isDeterministic() ? of(time, valueIfNonStochastic + value) : scalarOp(FmCuda.ADD, value);
	}

	@Override
	public RandomVariable sub(final double value) {
		return isDeterministic() ? of(time, valueIfNonStochastic - value) : scalarOp(FmCuda.SUB, value);
	}

	@Override
	public RandomVariable bus(final double value) {
		return isDeterministic() ? of(time, value - valueIfNonStochastic) : scalarOp(FmCuda.BUS, value);
	}

	@Override
	public RandomVariable mult(final double value) {
		return isDeterministic() ? of(time, valueIfNonStochastic * value) : scalarOp(FmCuda.MULT, value);
	}

	@Override
	public RandomVariable div(final double value) {
		return isDeterministic() ? of(time, valueIfNonStochastic / value) : scalarOp(FmCuda.DIV, value);
	}

	@Override
	public RandomVariable vid(final double value) {
		return isDeterministic() ? of(time, value / valueIfNonStochastic) : scalarOp(FmCuda.VID, value);
	}

	@Override
	public RandomVariable pow(final double exponent) {
		return isDeterministic() ? of(time, Math.pow(valueIfNonStochastic, exponent)) : scalarOp(FmCuda.POW, exponent);
	}

	@Override
	public RandomVariable average() {
		return of(time, getAverage());
	}

	/* ------------------------------------------------------------------ unary operations (RandomVariableCuda.java:1285-1387; RVF:866-954, 1287-1315, 1440-1451) */

	@Override
	public RandomVariable squared() {
		return isDeterministic() ? of(time, valueIfNonStochastic * valueIfNonStochastic) : unaryOp(FmCuda.SQUARED);
	}

	@Override
	public RandomVariable sqrt() {
		return isDeterministic() ? of(time, Math.sqrt(valueIfNonStochastic)) : unaryOp(FmCuda.SQRT);
	}

	@Override
	public RandomVariable invert() {
		return isDeterministic() ? of(time, 1.0 / valueIfNonStochastic) : unaryOp(FmCuda.INVERT);
	}

	@Override
	public RandomVariable abs() {
		return isDeterministic() ? of(time, Math.abs(valueIfNonStochastic)) : unaryOp(FmCuda.ABS);
	}

	@Override
	public RandomVariable exp() {
		return isDeterministic() ? of(time, Math.exp(valueIfNonStochastic)) : unaryOp(FmCuda.EXP);
	}

	@Override
	public RandomVariable log() {
		return isDeterministic() ? of(time, Math.log(valueIfNonStochastic)) : unaryOp(FmCuda.LOG);
	}

	@Override
	public RandomVariable sin() {
		return isDeterministic() ? of(time, Math.sin(valueIfNonStochastic)) : unaryOp(FmCuda.SIN);
	}

	@Override
	public RandomVariable cos() {
		return isDeterministic() ? of(time, Math.cos(valueIfNonStochastic)) : unaryOp(FmCuda.COS);
	}

	@Override
	public RandomVariable isNaN() {
		return isDeterministic() ? of(time, Double.isNaN(valueIfNonStochastic) ? 1.0 : 0.0) : unaryOp(FmCuda.ISNAN);
	}

	/* ------------------------------------------------------------------ operations with a random variable (RandomVariableCuda.java:1391-1579; RVF:960-1200)
	 * Dispatch, in this order: (1) an operand of higher type priority takes over with the mirrored method; (2) deterministic with
	 * deterministic in double on the host; (3) a deterministic operand enters as a scalar; (4) vector with vector on the device.
	 * The filtration time of a result is the maximum of the operands' (RVF:968; the reference GPU class forgets it in add/sub/bus). */

	@Override
	public RandomVariable add(final RandomVariable randomVariable) {
		if(randomVariable.getTypePriority() > this.getTypePriority()) {
			return randomVariable.add(this);
		}
		final double newTime = Math.max(time, randomVariable.getFiltrationTime());
		if(isDeterministic() && randomVariable.isDeterministic()) {
			return of(newTime, valueIfNonStochastic + randomVariable.doubleValue());
		}
		if(randomVariable.isDeterministic()) {
			return result(newTime, FmCuda.opVS(FmCuda.ADD, handle, randomVariable.doubleValue()));
		}
		final RandomVariableCuda other = onDevice(randomVariable);
		if(isDeterministic()) {
			return other.result(newTime, FmCuda.opVS(FmCuda.ADD, other.handle, valueIfNonStochastic));
		}
		return result(newTime, FmCuda.opVV(FmCuda.ADD, handle, other.handle));
	}

	@Override
	public RandomVariable sub(final RandomVariable randomVariable) {
		if(randomVariable.getTypePriority() > this.getTypePriority()) {
			return randomVariable.bus(this);
		}
		final double newTime = Math.max(time, randomVariable.getFiltrationTime());
		if(isDeterministic() && randomVariable.isDeterministic()) {
			return of(newTime, valueIfNonStochastic - randomVariable.doubleValue());
		}
		if(randomVariable.isDeterministic()) {
			return result(newTime, FmCuda.opVS(FmCuda.SUB, handle, randomVariable.doubleValue()));
		}
		final RandomVariableCuda other = onDevice(randomVariable);
		if(isDeterministic()) {
			return other.result(newTime, FmCuda.opVS(FmCuda.BUS, other.handle, valueIfNonStochastic));
		}
		return result(newTime, FmCuda.opVV(FmCuda.SUB, handle, other.handle));
	}

	@Override
	public RandomVariable bus(final RandomVariable randomVariable) {
		if(randomVariable.getTypePriority() > this.getTypePriority()) {
			return randomVariable.sub(this);
		}
		final double newTime = Math.max(time, randomVariable.getFiltrationTime());
		if(isDeterministic() && randomVariable.isDeterministic()) {
			return of(newTime, randomVariable.doubleValue() - valueIfNonStochastic);
		}
		if(randomVariable.isDeterministic()) {
			return result(newTime, FmCuda.opVS(FmCuda.BUS, handle, randomVariable.doubleValue()));
		}
		final RandomVariableCuda other = onDevice(randomVariable);
		if(isDeterministic()) {
			return other.result(newTime, FmCuda.opVS(FmCuda.SUB, other.handle, valueIfNonStochastic));
		}
		return result(newTime, FmCuda.opVV(FmCuda.BUS, handle, other.handle));
	}

	@Override
	public RandomVariable mult(final RandomVariable randomVariable) {
		if(randomVariable.getTypePriority() > this.getTypePriority()) {
			return randomVariable.mult(this);
		}
		final double newTime = Math.max(time, randomVariable.getFiltrationTime());
		if(isDeterministic() && randomVariable.isDeterministic()) {
			return of(newTime, valueIfNonStochastic * randomVariable.doubleValue());
		}
		if(randomVariable.isDeterministic()) {
			return result(newTime, FmCuda.opVS(FmCuda.MULT, handle, randomVariable.doubleValue()));
		}
		final RandomVariableCuda other = onDevice(randomVariable);
		if(isDeterministic()) {
			return other.result(newTime, FmCuda.opVS(FmCuda.MULT, other.handle, valueIfNonStochastic));
		}
		return result(newTime, FmCuda.opVV(FmCuda.MULT, handle, other.handle));
	}

	@Override
	public RandomVariable div(final RandomVariable randomVariable) {
		if(randomVariable.getTypePriority() > this.getTypePriority()) {
			return randomVariable.vid(this);
		}
		final double newTime = Math.max(time, randomVariable.getFiltrationTime());
		if(isDeterministic() && randomVariable.isDeterministic()) {
			return of(newTime, valueIfNonStochastic / randomVariable.doubleValue());
		}
		if(randomVariable.isDeterministic()) {
			return result(newTime, FmCuda.opVS(FmCuda.DIV, handle, randomVariable.doubleValue()));
		}
		final RandomVariableCuda other = onDevice(randomVariable);
		if(isDeterministic()) {
			return other.result(newTime, FmCuda.opVS(FmCuda.VID, other.handle, valueIfNonStochastic));
		}
		return result(newTime, FmCuda.opVV(FmCuda.DIV, handle, other.handle));
	}

	@Override
	public RandomVariable vid(final RandomVariable randomVariable) {
		if(randomVariable.getTypePriority() > this.getTypePriority()) {
			return randomVariable.div(this);													// RVF:1116-1119 (RandomVariableCuda.java:1513-1516 calls vid: defect)
		}
		final double newTime = Math.max(time, randomVariable.getFiltrationTime());
		if(isDeterministic() && randomVariable.isDeterministic()) {
			return of(newTime, randomVariable.doubleValue() / valueIfNonStochastic);
		}
		if(randomVariable.isDeterministic()) {
			return result(newTime, FmCuda.opVS(FmCuda.VID, handle, randomVariable.doubleValue()));
		}
		final RandomVariableCuda other = onDevice(randomVariable);
		if(isDeterministic()) {
			return other.result(newTime, FmCuda.opVS(FmCuda.DIV, other.handle, valueIfNonStochastic));
		}
		return result(newTime, FmCuda.opVV(FmCuda.VID, handle, other.handle));
	}

	@Override
	public RandomVariable cap(final RandomVariable randomVariable) {
		if(randomVariable.getTypePriority() > this.getTypePriority()) {
			return randomVariable.cap(this);
		}
		final double newTime = Math.max(time, randomVariable.getFiltrationTime());
		if(isDeterministic() && randomVariable.isDeterministic()) {
			return of(newTime, Math.min(valueIfNonStochastic, randomVariable.doubleValue()));
		}
		if(randomVariable.isDeterministic()) {
			return result(newTime, FmCuda.opVS(FmCuda.CAP, handle, randomVariable.doubleValue()));		// missing in RandomVariableCuda.java:1546-1555
		}
		final RandomVariableCuda other = onDevice(randomVariable);
		if(isDeterministic()) {
			return other.result(newTime, FmCuda.opVS(FmCuda.CAP, other.handle, valueIfNonStochastic));
		}
		return result(newTime, FmCuda.opVV(FmCuda.CAP, handle, other.handle));
	}

	@Override
	public RandomVariable floor(final RandomVariable randomVariable) {
		if(randomVariable.getTypePriority() > this.getTypePriority()) {
			return randomVariable.floor(this);
		}
		final double newTime = Math.max(time, randomVariable.getFiltrationTime());
		if(isDeterministic() && randomVariable.isDeterministic()) {
			return of(newTime, Math.max(valueIfNonStochastic, randomVariable.doubleValue()));
		}
		if(randomVariable.isDeterministic()) {
			return result(newTime, FmCuda.opVS(FmCuda.FLOOR, handle, randomVariable.doubleValue()));
		}
		final RandomVariableCuda other = onDevice(randomVariable);
		if(isDeterministic()) {
			return other.result(newTime, FmCuda.opVS(FmCuda.FLOOR, other.handle, valueIfNonStochastic));
		}
		return result(newTime, FmCuda.opVV(FmCuda.FLOOR, handle, other.handle));
	}

	/* ------------------------------------------------------------------ compound operations (RandomVariableCuda.java:1583-1695; RVF:1202-1438) */

	@Override
	public RandomVariable accrue(final RandomVariable rate, final double periodLength) {
		if(rate.getTypePriority() > this.getTypePriority()) {
			return rate.mult(periodLength).add(1.0).mult(this);									// RVF:1204-1207
		}
		final double newTime = Math.max(time, rate.getFiltrationTime());
		if(rate.isDeterministic()) {
			return mult(1.0 + rate.doubleValue() * periodLength);
		}
		if(isDeterministic()) {
			return rate.mult(periodLength).add(1.0).mult(valueIfNonStochastic);
		}
		final RandomVariableCuda r = onDevice(rate);
		return result(newTime, FmCuda.opVVS(FmCuda.ACCRUE, handle, r.handle, periodLength));
	}

	@Override
	public RandomVariable discount(final RandomVariable rate, final double periodLength) {
		if(rate.getTypePriority() > this.getTypePriority()) {
			return rate.mult(periodLength).add(1.0).vid(this);									// RVF:1232-1235 (RandomVariableCuda.java:1606 uses invert().mult: another rounding)
		}
		final double newTime = Math.max(time, rate.getFiltrationTime());
		if(rate.isDeterministic()) {
			return div(1.0 + rate.doubleValue() * periodLength);
		}
		if(isDeterministic()) {
			return rate.mult(periodLength).add(1.0).vid(valueIfNonStochastic);
		}
		final RandomVariableCuda r = onDevice(rate);
		return result(newTime, FmCuda.opVVS(FmCuda.DISCOUNT, handle, r.handle, periodLength));
	}

	@Override
	public RandomVariable choose(final RandomVariable valueIfTriggerNonNegative, final RandomVariable valueIfTriggerNegative) {
		if(isDeterministic()) {
			return valueIfNonStochastic >= 0 ? valueIfTriggerNonNegative : valueIfTriggerNegative;	// RVF:1264-1270
		}
		final double newTime = Math.max(time, Math.max(valueIfTriggerNonNegative.getFiltrationTime(), valueIfTriggerNegative.getFiltrationTime()));
		final RandomVariableCuda a = valueIfTriggerNonNegative.isDeterministic() ? null : onDevice(valueIfTriggerNonNegative);
		final RandomVariableCuda b = valueIfTriggerNegative.isDeterministic() ? null : onDevice(valueIfTriggerNegative);
		final long chosen = FmCuda.choose(handle,
				a != null ? a.handle : 0L, a != null ? 0.0 : valueIfTriggerNonNegative.doubleValue(),
				b != null ? b.handle : 0L, b != null ? 0.0 : valueIfTriggerNegative.doubleValue());
		return result(newTime, chosen);
	}

	@Override
	public RandomVariable addProduct(final RandomVariable factor1, final double factor2) {
		if(factor1.getTypePriority() > this.getTypePriority()) {
			return factor1.mult(factor2).add(this);												// RVF:1319-1322
		}
		final double newTime = Math.max(time, factor1.getFiltrationTime());
		if(factor1.isDeterministic()) {
			return add(factor1.doubleValue() * factor2);
		}
		if(isDeterministic()) {
			return factor1.mult(factor2).add(valueIfNonStochastic);
		}
		final RandomVariableCuda f1 = onDevice(factor1);
		return result(newTime, FmCuda.opVVS(FmCuda.ADDPRODUCT, handle, f1.handle, factor2));
	}

	@Override
	public RandomVariable addProduct(final RandomVariable factor1, final RandomVariable factor2) {
		if(factor1.getTypePriority() > this.getTypePriority() || factor2.getTypePriority() > this.getTypePriority()) {
			return factor1.mult(factor2).add(this);												// RVF:1355-1358
		}
		if(factor2.isDeterministic()) {
			return addProduct(factor1, factor2.doubleValue());
		}
		if(factor1.isDeterministic()) {
			return addProduct(factor2, factor1.doubleValue());
		}
		if(isDeterministic()) {
			return factor1.mult(factor2).add(valueIfNonStochastic);
		}
		final double newTime = Math.max(time, Math.max(factor1.getFiltrationTime(), factor2.getFiltrationTime()));
		final RandomVariableCuda f1 = onDevice(factor1), f2 = onDevice(factor2);
		return result(newTime, FmCuda.opVVV(FmCuda.ADDPRODUCT, handle, f1.handle, f2.handle));
	}

	@Override
	public RandomVariable addRatio(final RandomVariable numerator, final RandomVariable denominator) {
		if(numerator.getTypePriority() > this.getTypePriority() || denominator.getTypePriority() > this.getTypePriority()
				|| isDeterministic() || numerator.isDeterministic() || denominator.isDeterministic()) {
			return this.add(numerator.div(denominator));											// RVF:1396-1399
		}
		final double newTime = Math.max(time, Math.max(numerator.getFiltrationTime(), denominator.getFiltrationTime()));
		final RandomVariableCuda n = onDevice(numerator), d = onDevice(denominator);
		return result(newTime, FmCuda.opVVV(FmCuda.ADDRATIO, handle, n.handle, d.handle));
	}

	@Override
	public RandomVariable subRatio(final RandomVariable numerator, final RandomVariable denominator) {
		if(numerator.getTypePriority() > this.getTypePriority() || denominator.getTypePriority() > this.getTypePriority()
				|| isDeterministic() || numerator.isDeterministic() || denominator.isDeterministic()) {
			return this.sub(numerator.div(denominator));											// RVF:1419-1422
		}
		final double newTime = Math.max(time, Math.max(numerator.getFiltrationTime(), denominator.getFiltrationTime()));
		final RandomVariableCuda n = onDevice(numerator), d = onDevice(denominator);
		return result(newTime, FmCuda.opVVV(FmCuda.SUBRATIO, handle, n.handle, d.handle));
	}

	@Override
	public String toString() {
		return "RandomVariableCuda [time=" + time + ", size=" + size + (isDeterministic() ? ", value=" + valueIfNonStochastic : ", device vector 0x" + Long.toHexString(handle)) + "]";
	}
}
