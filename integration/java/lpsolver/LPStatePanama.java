// LPStatePanama.java — the same drop-in as LPStateNative, bound with the Foreign Function & Memory API
// (java.lang.foreign, final in JDK 22) instead of JNI: no C glue at all, liblps_b200.so is called directly.
//
// SOURCE COMPLETE, UNCOMPILED (no JDK in the build image).  Struct layouts mirror include/lps_b200.h and are
// pinned by static_asserts in csrc/lps_api.cu (sizeof(lps_options) == 64, sizeof(lps_run_result) == 64,
// sizeof(lps_objective_op) == 16) and by tests/test_abi.py on the ctypes side.
//
// The dense fast path: A, b, c live in off-heap MemorySegments (Arena.ofConfined / mapped files), so a caller
// that never needs BigDecimal[][] (LPInputReader -> doubles) pays no conversion and no copy before lps_load.
package lpsolver;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_FLOAT;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.StructLayout;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;

public final class LPStatePanama implements AutoCloseable {
  // ---- struct layouts (include/lps_b200.h) -------------------------------------------------------------
  /** lps_options: 64 bytes */
  static final StructLayout OPTIONS = MemoryLayout.structLayout(
      JAVA_DOUBLE.withName("epsilon"), JAVA_DOUBLE.withName("inf"), JAVA_INT.withName("device"),
      JAVA_INT.withName("time_kernels"), ADDRESS.withName("stream"), JAVA_INT.withName("update_variant"),
      JAVA_INT.withName("loop_mode"), JAVA_INT.withName("block_pivots"), JAVA_INT.withName("panel_ctas"),
      JAVA_INT.withName("pass_chunk_rows"), MemoryLayout.sequenceLayout(3, JAVA_INT).withName("reserved"));
  /** lps_run_result: 64 bytes */
  static final StructLayout RUN_RESULT = MemoryLayout.structLayout(
      JAVA_INT.withName("verdict"), JAVA_INT.withName("last_entering"), JAVA_INT.withName("last_leaving"),
      JAVA_INT.withName("pad_"), JAVA_LONG.withName("npivots"), JAVA_LONG.withName("total_pivots"),
      JAVA_DOUBLE.withName("v"), JAVA_FLOAT.withName("device_ms"), JAVA_FLOAT.withName("update_ms"),
      JAVA_LONG.withName("update_launches"), JAVA_LONG.withName("kernel_launches"));
  /** lps_objective_op: 16 bytes */
  static final StructLayout OBJECTIVE_OP = MemoryLayout.structLayout(
      JAVA_INT.withName("kind"), JAVA_INT.withName("index"), JAVA_DOUBLE.withName("coef"));

  // ---- downcall handles --------------------------------------------------------------------------------
  private static final Linker LINKER = Linker.nativeLinker();
  private static final SymbolLookup LIB = SymbolLookup.libraryLookup("liblps_b200.so", Arena.global());

  private static MethodHandle fn(String name, FunctionDescriptor d) {
    return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), d);
  }

  private static final MethodHandle DEFAULT_OPTIONS = fn("lps_default_options", FunctionDescriptor.ofVoid(ADDRESS));
  private static final MethodHandle CREATE = fn("lps_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  private static final MethodHandle DESTROY = fn("lps_destroy", FunctionDescriptor.of(JAVA_INT, ADDRESS));
  private static final MethodHandle LAST_ERROR = fn("lps_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
  private static final MethodHandle LOAD = fn("lps_load",
      FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, JAVA_DOUBLE));
  private static final MethodHandle LOAD_AUX = fn("lps_load_aux",
      FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS));
  private static final MethodHandle GET_ENTERING = fn("lps_get_entering", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  private static final MethodHandle GET_LEAVING = fn("lps_get_leaving", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
  private static final MethodHandle PIVOT = fn("lps_pivot", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT));
  private static final MethodHandle RUN = fn("lps_run", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS));
  private static final MethodHandle READ_V = fn("lps_read_v", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  private static final MethodHandle READ_B = fn("lps_read_b", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  private static final MethodHandle READ_C = fn("lps_read_c", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  private static final MethodHandle READ_ROW = fn("lps_read_row", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
  private static final MethodHandle READ_COL = fn("lps_read_col", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
  private static final MethodHandle READ_A = fn("lps_read_A", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));
  private static final MethodHandle READ_POSITIONS = fn("lps_read_positions", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  private static final MethodHandle POSITION_OF = fn("lps_position_of", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
  private static final MethodHandle READ_PRIMAL = fn("lps_read_primal", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
  private static final MethodHandle READ_PIVOT_LOG = fn("lps_read_pivot_log",
      FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS));
  private static final MethodHandle FIRST_NONZERO = fn("lps_first_nonzero_in_row",
      FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
  private static final MethodHandle DROP_COLUMN = fn("lps_drop_column", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
  private static final MethodHandle REBUILD_OBJECTIVE = fn("lps_rebuild_objective",
      FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT));

  // ---- state ---------------------------------------------------------------------------------------
  private final Arena arena = Arena.ofConfined();
  private MemorySegment handle = MemorySegment.NULL; // lps_handle
  int m, n;

  /** LPState(A, b, c, v, …, m, n) with the arrays already off-heap (row-major A, leading dimension lda). */
  public LPStatePanama(MemorySegment a, long lda, MemorySegment b, MemorySegment c, double v, int m, int n,
      double epsilon, double inf) {
    this.m = m;
    this.n = n;
    create(epsilon, inf);
    check(invokeInt(LOAD, handle, m, n, a, lda, b, c, v));
  }

  /** LPSolver.convertIntoAuxLP (LPSolver.java:283-321) on the device. */
  public static LPStatePanama auxiliary(MemorySegment a, long lda, MemorySegment b, int m, int n, double epsilon,
      double inf) {
    LPStatePanama s = new LPStatePanama(m, n + 1);
    s.create(epsilon, inf);
    s.check(invokeInt(LOAD_AUX, s.handle, m, n, a, lda, b));
    return s;
  }

  private LPStatePanama(int m, int n) {
    this.m = m;
    this.n = n;
  }

  private void create(double epsilon, double inf) {
    MemorySegment opts = arena.allocate(OPTIONS);
    try {
      DEFAULT_OPTIONS.invokeExact(opts);
    } catch (Throwable t) {
      throw new IllegalStateException(t);
    }
    opts.set(JAVA_DOUBLE, 0, epsilon); // LPState.DEF_EPSILON, LPState.java:20
    opts.set(JAVA_DOUBLE, 8, inf); //     LPState.DEF_INF,     LPState.java:21
    MemorySegment out = arena.allocate(ADDRESS);
    int rc = invokeInt(CREATE, out, opts);
    if (rc != 0) throw new IllegalStateException("lps_create failed (" + rc + "): no usable CUDA device, no CPU fallback");
    handle = out.get(ADDRESS, 0);
  }

  // ---- the reference's method set ----------------------------------------------------------------------
  /** LPState.getEntering — LPState.java:274-285 */
  public int getEntering() {
    MemorySegment e = arena.allocate(JAVA_INT);
    check(invokeInt(GET_ENTERING, handle, e));
    return e.get(JAVA_INT, 0);
  }

  /** LPState.getLeaving — LPState.java:287-305 */
  public int getLeaving(int entering) {
    MemorySegment l = arena.allocate(JAVA_INT);
    check(invokeInt(GET_LEAVING, handle, entering, l));
    return l.get(JAVA_INT, 0);
  }

  /** LPState.pivot — LPState.java:114-181 */
  public void pivot(int entering, int leaving) throws SolutionException {
    int rc = invokeInt(PIVOT, handle, entering, leaving);
    if (rc == -1) throw new IllegalArgumentException(lastError());
    if (rc != 0) throw new SolutionException(lastError());
  }

  /** the loops of LPSolver.simplex (LPSolver.java:101-112) / solveAuxLP (:141-161); returns the lps_run_result */
  public MemorySegment run(long maxPivots) throws SolutionException {
    MemorySegment r = arena.allocate(RUN_RESULT);
    int rc = invokeInt(RUN, handle, maxPivots, r);
    if (rc != 0) throw new SolutionException(lastError());
    return r; // verdict at offset 0, npivots at 16, v at 32
  }

  public double v() {
    MemorySegment x = arena.allocate(JAVA_DOUBLE);
    check(invokeInt(READ_V, handle, x));
    return x.get(JAVA_DOUBLE, 0);
  }

  public double[] b() { return readDoubles(READ_B, m); }

  public double[] c() { return readDoubles(READ_C, n); }

  public double[] row(int i) {
    MemorySegment x = arena.allocate(JAVA_DOUBLE, n);
    check(invokeInt(READ_ROW, handle, i, x));
    return x.toArray(JAVA_DOUBLE);
  }

  public double[] column(int j) {
    MemorySegment x = arena.allocate(JAVA_DOUBLE, m);
    check(invokeInt(READ_COL, handle, j, x));
    return x.toArray(JAVA_DOUBLE);
  }

  /** the whole (pivoted) A into an off-heap segment, for callers that want the reference's aliasing back */
  public void readA(MemorySegment dst, long lda) { check(invokeInt(READ_A, handle, dst, lda)); }

  public int[] positions() {
    MemorySegment p = arena.allocate(JAVA_INT, m + n);
    check(invokeInt(READ_POSITIONS, handle, p));
    return p.toArray(JAVA_INT);
  }

  public int positionOf(int var) {
    MemorySegment p = arena.allocate(JAVA_INT);
    check(invokeInt(POSITION_OF, handle, var, p));
    return p.get(JAVA_INT, 0);
  }

  public double[] readPrimal(int nvars) {
    MemorySegment x = arena.allocate(JAVA_DOUBLE, Math.max(1, nvars));
    check(invokeInt(READ_PRIMAL, handle, nvars, x));
    return x.asSlice(0, 8L * nvars).toArray(JAVA_DOUBLE);
  }

  public int[] pivotLog() {
    MemorySegment cnt = arena.allocate(JAVA_LONG);
    check(invokeInt(READ_PIVOT_LOG, handle, MemorySegment.NULL, 0L, cnt));
    long k = cnt.get(JAVA_LONG, 0);
    MemorySegment pairs = arena.allocate(JAVA_INT, Math.max(1, 2 * k));
    if (k > 0) check(invokeInt(READ_PIVOT_LOG, handle, pairs, k, cnt));
    return pairs.asSlice(0, 8 * k).toArray(JAVA_INT);
  }

  // ---- phase-1 support (LPSolver.java:166-246) ---------------------------------------------------------
  public int firstNonzeroInRow(int row) {
    MemorySegment j = arena.allocate(JAVA_INT);
    check(invokeInt(FIRST_NONZERO, handle, row, j));
    return j.get(JAVA_INT, 0);
  }

  public void dropColumn(int j) {
    check(invokeInt(DROP_COLUMN, handle, j));
    n -= 1;
  }

  /** ops: {kind, index, coef} triples in the order LPSolver.restoreInitialLP iterates (LPSolver.java:213-233) */
  public void rebuildObjective(int[] kinds, int[] indexes, double[] coefs) {
    MemorySegment ops = arena.allocate(OBJECTIVE_OP, Math.max(1, kinds.length));
    for (int k = 0; k < kinds.length; k++) {
      ops.set(JAVA_INT, 16L * k, kinds[k]);
      ops.set(JAVA_INT, 16L * k + 4, indexes[k]);
      ops.set(JAVA_DOUBLE, 16L * k + 8, coefs[k]);
    }
    check(invokeInt(REBUILD_OBJECTIVE, handle, ops, kinds.length));
  }

  @Override
  public void close() {
    if (!handle.equals(MemorySegment.NULL)) {
      invokeInt(DESTROY, handle);
      handle = MemorySegment.NULL;
    }
    arena.close();
  }

  // ---- helpers ---------------------------------------------------------------------------------------
  private double[] readDoubles(MethodHandle h, int len) {
    MemorySegment x = arena.allocate(JAVA_DOUBLE, Math.max(1, len));
    check(invokeInt(h, handle, x));
    return x.asSlice(0, 8L * len).toArray(JAVA_DOUBLE);
  }

  private String lastError() {
    try {
      MemorySegment s = (MemorySegment) LAST_ERROR.invoke(handle);
      return s.reinterpret(4096).getString(0);
    } catch (Throwable t) {
      return "lps_last_error failed: " + t;
    }
  }

  private void check(int rc) {
    if (rc == 0) return;
    if (rc == -1) throw new IllegalArgumentException(lastError()); // Validate.isTrue, LPState.java:288
    throw new IllegalStateException("liblps_b200: " + lastError() + " (status " + rc + ")");
  }

  private static int invokeInt(MethodHandle h, Object... args) {
    try {
      return (int) h.invokeWithArguments(args);
    } catch (Throwable t) {
      throw new IllegalStateException(t);
    }
  }
}
