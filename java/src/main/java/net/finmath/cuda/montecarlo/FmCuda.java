/*
 * Bridge from the Java shim to the B200-native runtime (include/fmcuda.h -> libfmcuda.so) over java.lang.foreign (JDK 22+).
 *
 * Replaces everything RandomVariableCuda.java:8-16 imports from jcuda (cuInit/cuCtxCreate/cuModuleLoad/cuMemAlloc/cuMemcpyHtoD/
 * cuMemcpyDtoH/cuLaunchKernel, RandomVariableCuda.java:119-558) and jcurand (BrownianMotionCudaWithRandomVariableCuda.java:141-182).
 * Written against include/fmcuda.h of this repository; no JVM exists in the build environment of this repository, so these
 * sources are delivered uncompiled (see java/README.md); the same ABI is exercised by the ctypes binding finmath_cuda/_capi.py
 * and the C++ mirror include/finmath/RandomVariableCuda.hpp.
 */
package net.finmath.cuda.montecarlo;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_FLOAT;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

import java.io.IOException;
import java.io.InputStream;
import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;
import java.lang.ref.Cleaner;
import java.nio.file.Files;
import java.nio.file.Path;
import java.nio.file.StandardCopyOption;

/**
 * Thin, stateless bridge to <code>libfmcuda.so</code>. All methods are thread safe (the native runtime takes its own lock and binds
 * the calling thread to its device; there is no context-owning executor thread as in RandomVariableCuda.java:155, 210-215).
 */
final class FmCuda {

	/* opcodes of fmc_op_* (include/fmcuda.h) */
	static final int CAP = 1, FLOOR = 2, ADD = 3, SUB = 4, BUS = 5, MULT = 6, DIV = 7, VID = 8, POW = 9;
	static final int SQUARED = 20, SQRT = 21, EXP = 22, LOG = 23, SIN = 24, COS = 25, INVERT = 26, ABS = 27, ISNAN = 28;
	static final int ACCRUE = 40, DISCOUNT = 41, ADDPRODUCT = 42, CHOOSE = 43, ADDRATIO = 44, SUBRATIO = 45;
	/* reduction kinds of fmc_reduce */
	static final int RED_SUM = 1, RED_AVERAGE = 2, RED_VARIANCE = 3, RED_SAMPLE_VARIANCE = 4, RED_MIN = 5, RED_MAX = 6, RED_AVERAGE_W = 7, RED_VARIANCE_W = 8;
	/* status codes */
	static final int ERR_INVALID = -1, ERR_OOM = -2, ERR_CUDA = -3, ERR_SIZE = -4, ERR_NOT_INIT = -5, ERR_COMM = -6, ERR_UNSUPPORTED = -7;

	private static final Linker LINKER = Linker.nativeLinker();
	private static final SymbolLookup LIB = SymbolLookup.libraryLookup(locateLibrary(), Arena.global());

	/** One cleaner for all device vectors: replaces the ReferenceQueue recycling of RandomVariableCuda.java:295-306. */
	static final Cleaner CLEANER = Cleaner.create();

	private static final MethodHandle INIT = h("fmc_init", FunctionDescriptor.of(JAVA_INT, JAVA_INT));
	private static final MethodHandle LAST_ERROR = h("fmc_last_error", FunctionDescriptor.of(ADDRESS));
	private static final MethodHandle FROM_F64 = h("fmc_vec_from_f64", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS));
	private static final MethodHandle FROM_F32 = h("fmc_vec_from_f32", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS));
	private static final MethodHandle TO_F64 = h("fmc_vec_to_f64", FunctionDescriptor.of(JAVA_INT, JAVA_LONG, ADDRESS, JAVA_LONG));
	private static final MethodHandle TO_F32 = h("fmc_vec_to_f32", FunctionDescriptor.of(JAVA_INT, JAVA_LONG, ADDRESS, JAVA_LONG));
	private static final MethodHandle GET = h("fmc_vec_get", FunctionDescriptor.of(JAVA_INT, JAVA_LONG, JAVA_LONG, ADDRESS));
	private static final MethodHandle RELEASE = h("fmc_vec_release", FunctionDescriptor.of(JAVA_INT, JAVA_LONG));
	private static final MethodHandle OP_VS = h("fmc_op_vs", FunctionDescriptor.of(JAVA_INT, JAVA_INT, JAVA_LONG, JAVA_DOUBLE, ADDRESS));
	private static final MethodHandle OP_V = h("fmc_op_v", FunctionDescriptor.of(JAVA_INT, JAVA_INT, JAVA_LONG, ADDRESS));
	private static final MethodHandle OP_VV = h("fmc_op_vv", FunctionDescriptor.of(JAVA_INT, JAVA_INT, JAVA_LONG, JAVA_LONG, ADDRESS));
	private static final MethodHandle OP_VVS = h("fmc_op_vvs", FunctionDescriptor.of(JAVA_INT, JAVA_INT, JAVA_LONG, JAVA_LONG, JAVA_DOUBLE, ADDRESS));
	private static final MethodHandle OP_VVV = h("fmc_op_vvv", FunctionDescriptor.of(JAVA_INT, JAVA_INT, JAVA_LONG, JAVA_LONG, JAVA_LONG, ADDRESS));
	private static final MethodHandle OP_CHOOSE = h("fmc_op_choose", FunctionDescriptor.of(JAVA_INT, JAVA_LONG, JAVA_LONG, JAVA_DOUBLE, JAVA_LONG, JAVA_DOUBLE, ADDRESS));
	private static final MethodHandle REDUCE = h("fmc_reduce", FunctionDescriptor.of(JAVA_INT, JAVA_INT, JAVA_LONG, JAVA_LONG, ADDRESS));
	private static final MethodHandle QUANTILE = h("fmc_quantile", FunctionDescriptor.of(JAVA_INT, JAVA_LONG, JAVA_DOUBLE, ADDRESS));
	private static final MethodHandle QUANTILE_EXPECTATION = h("fmc_quantile_expectation", FunctionDescriptor.of(JAVA_INT, JAVA_LONG, JAVA_DOUBLE, JAVA_DOUBLE, ADDRESS));
	private static final MethodHandle HISTOGRAM = h("fmc_histogram", FunctionDescriptor.of(JAVA_INT, JAVA_LONG, ADDRESS, JAVA_INT, ADDRESS));
	private static final MethodHandle REGRESSION = h("fmc_regression_normal_eq", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_LONG, ADDRESS, ADDRESS));
	private static final MethodHandle BROWNIAN = h("fmc_brownian_generate", FunctionDescriptor.of(JAVA_INT, JAVA_INT, JAVA_LONG, JAVA_INT, JAVA_INT, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS));
	private static final MethodHandle SYNC = h("fmc_sync", FunctionDescriptor.of(JAVA_INT));
	private static final MethodHandle POOL_TRIM = h("fmc_pool_trim", FunctionDescriptor.of(JAVA_INT));
	private static final MethodHandle POOL_PURGE = h("fmc_pool_purge", FunctionDescriptor.of(JAVA_INT));
	private static final MethodHandle SET_OPTION = h("fmc_set_option", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_DOUBLE));

	static {
		// the same system property the reference reads (RandomVariableCuda.java:161, 177); -1 = last device
		check(invokeInt(INIT, Integer.getInteger("net.finmath.montecarlo.opencl.RandomVariableCuda.deviceIndex", -1)));
	}

	private FmCuda() {}

	private static MethodHandle h(final String name, final FunctionDescriptor descriptor) {
		return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError("libfmcuda.so lacks " + name)), descriptor);
	}

	/**
	 * The native library: the path given by -Dnet.finmath.cuda.library, else the class-path resource
	 * net/finmath/cuda/montecarlo/libfmcuda.so copied to a temporary file (the reference ships its kernel source as a class-path
	 * resource the same way, pom.xml:321-326 and RandomVariableCuda.java:181-206), else the loader's search path.
	 */
	private static String locateLibrary() {
		final String configured = System.getProperty("net.finmath.cuda.library");
		if(configured != null) {
			return configured;
		}
		try(InputStream in = FmCuda.class.getResourceAsStream("libfmcuda.so")) {
			if(in != null) {
				final Path tmp = Files.createTempFile("libfmcuda", ".so");
				tmp.toFile().deleteOnExit();
				Files.copy(in, tmp, StandardCopyOption.REPLACE_EXISTING);
				return tmp.toString();
			}
		} catch(final IOException e) {
			throw new UnsatisfiedLinkError("cannot unpack libfmcuda.so: " + e.getMessage());
		}
		return "libfmcuda.so";
	}

	private static int invokeInt(final MethodHandle handle, final Object... arguments) {
		try {
			return (int) handle.invokeWithArguments(arguments);
		} catch(final RuntimeException | Error e) {
			throw e;
		} catch(final Throwable t) {
			throw new RuntimeException(t);
		}
	}

	/** Status code to exception, like JCuda's exceptions and RandomVariableCuda.java:373-376, 464, 477-479. */
	static void check(final int status) {
		if(status == 0) {
			return;
		}
		String message;
		try {
			message = ((MemorySegment) LAST_ERROR.invokeExact()).reinterpret(1024).getString(0);
		} catch(final Throwable t) {
			message = "fmcuda error " + status;
		}
		switch(status) {
		case ERR_OOM: throw new OutOfMemoryError(message);
		case ERR_SIZE: throw new ArrayIndexOutOfBoundsException(message);
		case ERR_INVALID: throw new IllegalArgumentException(message);
		case ERR_UNSUPPORTED: throw new UnsupportedOperationException(message);
		default: throw new RuntimeException(message);
		}
	}

	/* ---- vectors ---- */

	/** createRandomVariable(time, double[]): the (float) cast of RandomVariableCuda.java:768-774 happens in the runtime. */
	static long upload(final double[] values) {
		try(Arena arena = Arena.ofConfined()) {
			final MemorySegment host = arena.allocateFrom(JAVA_DOUBLE, values);
			final MemorySegment out = arena.allocate(JAVA_LONG);
			check(invokeInt(FROM_F64, host, (long) values.length, out));
			return out.get(JAVA_LONG, 0);
		}
	}

	static long upload(final float[] values) {
		try(Arena arena = Arena.ofConfined()) {
			final MemorySegment host = arena.allocateFrom(JAVA_FLOAT, values);
			final MemorySegment out = arena.allocate(JAVA_LONG);
			check(invokeInt(FROM_F32, host, (long) values.length, out));
			return out.get(JAVA_LONG, 0);
		}
	}

	/** getRealizations(): RandomVariableCuda.java:1115-1122 (download through pinned staging, widened to double). */
	static double[] downloadAsDouble(final long vector, final long size) {
		try(Arena arena = Arena.ofConfined()) {
			final MemorySegment host = arena.allocate(JAVA_DOUBLE, size);
			check(invokeInt(TO_F64, vector, host, size));
			return host.toArray(JAVA_DOUBLE);
		}
	}

	static float[] downloadAsFloat(final long vector, final long size) {
		try(Arena arena = Arena.ofConfined()) {
			final MemorySegment host = arena.allocate(JAVA_FLOAT, size);
			check(invokeInt(TO_F32, vector, host, size));
			return host.toArray(JAVA_FLOAT);
		}
	}

	static double get(final long vector, final long index) {
		try(Arena arena = Arena.ofConfined()) {
			final MemorySegment out = arena.allocate(JAVA_DOUBLE);
			check(invokeInt(GET, vector, index, out));
			return out.get(JAVA_DOUBLE, 0);
		}
	}

	/** Called by the Cleaner when a RandomVariableCuda has become unreachable. Errors are swallowed: the runtime may be gone at JVM exit. */
	static void release(final long vector) {
		try {
			invokeInt(RELEASE, vector);
		} catch(final Throwable t) {
			// nothing sensible can be done on the cleaner thread
		}
	}

	/* ---- recorded operations (nothing runs until a value is demanded) ---- */

	static long opVS(final int opcode, final long x, final double s) {
		try(Arena arena = Arena.ofConfined()) {
			final MemorySegment out = arena.allocate(JAVA_LONG);
			check(invokeInt(OP_VS, opcode, x, s, out));
			return out.get(JAVA_LONG, 0);
		}
	}

	static long opV(final int opcode, final long x) {
		try(Arena arena = Arena.ofConfined()) {
			final MemorySegment out = arena.allocate(JAVA_LONG);
			check(invokeInt(OP_V, opcode, x, out));
			return out.get(JAVA_LONG, 0);
		}
	}

	static long opVV(final int opcode, final long x, final long y) {
		try(Arena arena = Arena.ofConfined()) {
			final MemorySegment out = arena.allocate(JAVA_LONG);
			check(invokeInt(OP_VV, opcode, x, y, out));
			return out.get(JAVA_LONG, 0);
		}
	}

	static long opVVS(final int opcode, final long x, final long y, final double s) {
		try(Arena arena = Arena.ofConfined()) {
			final MemorySegment out = arena.allocate(JAVA_LONG);
			check(invokeInt(OP_VVS, opcode, x, y, s, out));
			return out.get(JAVA_LONG, 0);
		}
	}

	static long opVVV(final int opcode, final long x, final long y, final long z) {
		try(Arena arena = Arena.ofConfined()) {
			final MemorySegment out = arena.allocate(JAVA_LONG);
			check(invokeInt(OP_VVV, opcode, x, y, z, out));
			return out.get(JAVA_LONG, 0);
		}
	}

	/** trigger >= 0 ? a : b, where a handle of 0 selects the scalar of that branch. */
	static long choose(final long trigger, final long ifNonNegative, final double scalarNonNegative, final long ifNegative, final double scalarNegative) {
		try(Arena arena = Arena.ofConfined()) {
			final MemorySegment out = arena.allocate(JAVA_LONG);
			check(invokeInt(OP_CHOOSE, trigger, ifNonNegative, scalarNonNegative, ifNegative, scalarNegative, out));
			return out.get(JAVA_LONG, 0);
		}
	}

	/* ---- reductions and order statistics (these demand values: the pending chain of x is fused into the reduction kernel) ---- */

	static double reduce(final int kind, final long x, final long weightsOrZero) {
		try(Arena arena = Arena.ofConfined()) {
			final MemorySegment out = arena.allocate(JAVA_DOUBLE);
			check(invokeInt(REDUCE, kind, x, weightsOrZero, out));
			return out.get(JAVA_DOUBLE, 0);
		}
	}

	static double quantile(final long x, final double quantile) {
		try(Arena arena = Arena.ofConfined()) {
			final MemorySegment out = arena.allocate(JAVA_DOUBLE);
			check(invokeInt(QUANTILE, x, quantile, out));
			return out.get(JAVA_DOUBLE, 0);
		}
	}

	static double quantileExpectation(final long x, final double quantileStart, final double quantileEnd) {
		try(Arena arena = Arena.ofConfined()) {
			final MemorySegment out = arena.allocate(JAVA_DOUBLE);
			check(invokeInt(QUANTILE_EXPECTATION, x, quantileStart, quantileEnd, out));
			return out.get(JAVA_DOUBLE, 0);
		}
	}

	static double[] histogram(final long x, final double[] intervalPoints) {
		try(Arena arena = Arena.ofConfined()) {
			final MemorySegment points = arena.allocateFrom(JAVA_DOUBLE, intervalPoints);
			final MemorySegment out = arena.allocate(JAVA_DOUBLE, intervalPoints.length + 1L);
			check(invokeInt(HISTOGRAM, x, points, intervalPoints.length, out));
			return out.toArray(JAVA_DOUBLE);
		}
	}

	/**
	 * Normal equations of a conditional-expectation regression in one fused pass: XtX[i*k+j] = average(b_i b_j), Xty[i] = average(y b_i).
	 * A basis handle of 0 selects the constant scalars[i].
	 */
	static void regressionNormalEquations(final long[] basis, final double[] scalars, final long y, final double[] xtx, final double[] xty) {
		final int k = basis.length;
		try(Arena arena = Arena.ofConfined()) {
			final MemorySegment b = arena.allocateFrom(JAVA_LONG, basis);
			final MemorySegment s = arena.allocateFrom(JAVA_DOUBLE, scalars);
			final MemorySegment a = arena.allocate(JAVA_DOUBLE, (long) k * k);
			final MemorySegment c = arena.allocate(JAVA_DOUBLE, k);
			check(invokeInt(REGRESSION, b, s, k, y, a, c));
			MemorySegment.copy(a, JAVA_DOUBLE, 0, xtx, 0, k * k);
			MemorySegment.copy(c, JAVA_DOUBLE, 0, xty, 0, k);
		}
	}

	/**
	 * Brownian increments of paths [pathStart, pathEnd) of a numberOfPaths-path motion: the MT19937 stream of
	 * BrownianMotionFromMersenneRandomNumbers (commons-math3), jump-ahead per GPU block, AS241 inverse normal. Returns T*F handles, index t*F+f.
	 */
	static long[] brownianIncrements(final int seedMode, final long seed, final int numberOfTimeSteps, final int numberOfFactors, final long pathStart, final long pathEnd, final double[] sqrtOfTimeSteps) {
		try(Arena arena = Arena.ofConfined()) {
			final MemorySegment sqrtDt = arena.allocateFrom(JAVA_DOUBLE, sqrtOfTimeSteps);
			final MemorySegment out = arena.allocate(JAVA_LONG, (long) numberOfTimeSteps * numberOfFactors);
			check(invokeInt(BROWNIAN, seedMode, seed, numberOfTimeSteps, numberOfFactors, pathStart, pathEnd, sqrtDt, out));
			return out.toArray(JAVA_LONG);
		}
	}

	static void synchronize() { check(invokeInt(SYNC)); }
	/** RandomVariableCuda.clean(), RandomVariableCuda.java:751-753. */
	static void poolTrim() { check(invokeInt(POOL_TRIM)); }
	/** RandomVariableCuda.purge(), RandomVariableCuda.java:755-757. */
	static void poolPurge() { check(invokeInt(POOL_PURGE)); }

	static void setOption(final String key, final double value) {
		try(Arena arena = Arena.ofConfined()) {
			check(invokeInt(SET_OPTION, arena.allocateFrom(key), value));
		}
	}
}
