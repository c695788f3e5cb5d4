// LPStateNative.java — drop-in body for lpsolver.LPState (LPState.java:17-320) over liblps_b200.so via JNI.
//
// SOURCE COMPLETE, UNCOMPILED: the build image has no JDK.  The C side of every native method below is
// integration/java/jni/lps_b200_jni.c, which IS compiled and run in this repository's test-suite against a
// stand-in jni.h (tests/jni_mock/jni.h, tests/test_jni_glue.py): same functions, same call order, borrowed
// primitive-array buffers as GetPrimitiveArrayCritical hands them out.
//
// Same package, same method set and field meaning as the reference class:
//   getEntering()            LPState.java:274-285
//   getLeaving(int)          LPState.java:287-305   (IllegalArgumentException for an out-of-range index, :288)
//   pivot(int,int)           LPState.java:114-181   (pivotSequentially / pivotConcurrently give the same values)
//   exchangeIndexes          LPState.java:311-320   (kept on the device as a position permutation)
// plus the loops of LPSolver.simplex / solveAuxLP as ONE call (run), and the device-side helpers that
// LPSolver.handleInitialization / performDegeneratePivot / restoreInitialLP need (LPSolver.java:166-246).
package lpsolver;

import java.math.BigDecimal;
import java.math.MathContext;
import java.util.HashMap;

public class LPStateNative implements AutoCloseable {
  static {
    System.loadLibrary("lps_b200_jni"); // JNI glue, linked against liblps_b200.so
  }

  public static final BigDecimal DEF_EPSILON = LPState.DEF_EPSILON; // 1e-9   LPState.java:20
  public static final BigDecimal DEF_INF = LPState.DEF_INF; //         1e50   LPState.java:21

  /** verdicts of {@link #run(long)} (lps_verdict in lps_b200.h) */
  public static final int RUNNING = 0, OPTIMAL = 1, UNBOUNDED = 2, PIVOT_CAP = 3;

  /** result of {@link #run(long)}: lps_run_result */
  public static final class RunResult {
    public int verdict, lastEntering, lastLeaving;
    public long pivots, totalPivots;
    public double v;
    public float deviceMs;
  }

  private long handle; // lps_handle
  int m, n; // same names as the reference's package-private fields
  HashMap<Integer, String> variables; // position -> name   (LPState.java:27)
  HashMap<String, Integer> coefficients; // name -> position (LPState.java:28)
  private String[] names0; // name of variable id k (ids = positions at load time)

  // ---- constructors: LPState.java:37-112 ---------------------------------------------------------------
  /** LPState(A, b, c, m, n) — LPState.java:88-99.  The BigDecimal arrays are converted with doubleValue(). */
  LPStateNative(BigDecimal[][] A, BigDecimal[] b, BigDecimal[] c, int m, int n) {
    this(flatten(A, m, n), toDoubles(b, m), toDoubles(c, n), 0.0, null, null, m, n,
        DEF_EPSILON.doubleValue(), DEF_INF.doubleValue());
  }

  /** LPState(A, b, c, variables, coefficients, m, n) — LPState.java:61-72 */
  LPStateNative(BigDecimal[][] A, BigDecimal[] b, BigDecimal[] c, HashMap<Integer, String> variables,
      HashMap<String, Integer> coefficients, int m, int n) {
    this(flatten(A, m, n), toDoubles(b, m), toDoubles(c, n), 0.0, variables, coefficients, m, n,
        DEF_EPSILON.doubleValue(), DEF_INF.doubleValue());
  }

  /** LPState(A, b, c, v, variables, coefficients, m, n) — LPState.java:101-112 */
  LPStateNative(BigDecimal[][] A, BigDecimal[] b, BigDecimal[] c, BigDecimal v,
      HashMap<Integer, String> variables, HashMap<String, Integer> coefficients, int m, int n) {
    this(flatten(A, m, n), toDoubles(b, m), toDoubles(c, n), v.doubleValue(), variables, coefficients, m, n,
        DEF_EPSILON.doubleValue(), DEF_INF.doubleValue());
  }

  /** LPState(A, b, c, v, variables, coefficients, m, n, printRounder, rounder, epsilon, INF) — LPState.java:37-59.
   *  The MathContext arguments have no binary64 analogue (device arithmetic is IEEE round-to-nearest-even);
   *  epsilon and INF are honoured. */
  LPStateNative(BigDecimal[][] A, BigDecimal[] b, BigDecimal[] c, BigDecimal v,
      HashMap<Integer, String> variables, HashMap<String, Integer> coefficients, int m, int n,
      MathContext printRounder, MathContext rounder, BigDecimal epsilon, BigDecimal INF) {
    this(flatten(A, m, n), toDoubles(b, m), toDoubles(c, n), v.doubleValue(), variables, coefficients, m, n,
        epsilon.doubleValue(), INF.doubleValue());
  }

  /** Dense fast path: row-major double[] (no BigDecimal objects at all). */
  public LPStateNative(double[] aRowMajor, double[] b, double[] c, double v,
      HashMap<Integer, String> variables, HashMap<String, Integer> coefficients, int m, int n,
      double epsilon, double inf) {
    this.m = m;
    this.n = n;
    this.variables = variables;
    this.coefficients = coefficients;
    rememberNames();
    handle = nCreate(epsilon, inf, -1);
    if (handle == 0) throw new IllegalStateException("lps_create failed: no usable CUDA device (there is no CPU fallback)");
    check(nLoad(handle, m, n, aRowMajor, n, b, c, v));
  }

  /** LPSolver.convertIntoAuxLP (LPSolver.java:283-321) built on the device: column n = -1, objective (0,…,0,-1). */
  public static LPStateNative auxiliary(double[] aRowMajor, double[] b, int m, int n,
      HashMap<Integer, String> variables, HashMap<String, Integer> coefficients, double epsilon, double inf) {
    LPStateNative s = new LPStateNative();
    s.m = m;
    s.n = n + 1;
    s.variables = variables;
    s.coefficients = coefficients;
    s.rememberNames();
    s.handle = nCreate(epsilon, inf, -1);
    if (s.handle == 0) throw new IllegalStateException("lps_create failed: no usable CUDA device");
    s.check(nLoadAux(s.handle, m, n, aRowMajor, n, b));
    return s;
  }

  private LPStateNative() {}

  // ---- the reference's method set ----------------------------------------------------------------------
  /** LPState.getEntering — LPState.java:274-285: first i with c[i] > epsilon, else -1. */
  public int getEntering() {
    int[] e = new int[1];
    check(nGetEntering(handle, e));
    return e[0];
  }

  /** LPState.getLeaving — LPState.java:287-305. */
  public int getLeaving(int entering) {
    int[] l = new int[1];
    check(nGetLeaving(handle, entering, l));
    return l[0];
  }

  /** LPState.pivot — LPState.java:114-181 (+ exchangeIndexes :311-320). */
  public void pivot(int entering, int leaving) throws SolutionException {
    int rc = nPivot(handle, entering, leaving);
    if (rc == -1) throw new IllegalArgumentException(nLastError(handle));
    if (rc != 0) throw new SolutionException(nLastError(handle));
    refreshNames();
  }

  /** The whole loop of LPSolver.simplex (LPSolver.java:101-112) / solveAuxLP (:141-161) on the device. */
  public RunResult run(long maxPivots) throws SolutionException {
    RunResult r = new RunResult();
    long[] l = new long[5];
    double[] d = new double[2];
    int rc = nRun(handle, maxPivots, l, d);
    if (rc != 0) throw new SolutionException(nLastError(handle));
    r.verdict = (int) l[0];
    r.lastEntering = (int) l[1];
    r.lastLeaving = (int) l[2];
    r.pivots = l[3];
    r.totalPivots = l[4];
    r.v = d[0];
    r.deviceMs = (float) d[1];
    refreshNames();
    return r;
  }

  // ---- fields LPSolver reads (LPSolver.java:113,129,137,170,185,203-245) -----------------------------------
  public double v() {
    double[] x = new double[1];
    check(nReadV(handle, x));
    return x[0];
  }

  public double[] b() {
    double[] x = new double[m];
    check(nReadB(handle, x));
    return x;
  }

  public double[] c() {
    double[] x = new double[n];
    check(nReadC(handle, x));
    return x;
  }

  public double[] row(int i) {
    double[] x = new double[n];
    check(nReadRow(handle, i, x));
    return x;
  }

  public double[] column(int j) {
    double[] x = new double[m];
    check(nReadCol(handle, j, x));
    return x;
  }

  /** position -> variable id (the two name maps as a permutation). */
  public int[] positions() {
    int[] p = new int[m + n];
    check(nReadPositions(handle, p));
    return p;
  }

  /** where variable `var` sits now: &lt; n non-basic column, &gt;= n basic row (x0 tracking, LPSolver.java:151-155). */
  public int positionOf(int var) {
    int[] p = new int[1];
    check(nPositionOf(handle, var, p));
    return p[0];
  }

  /** primal values of variables 0..nvars-1 (io_files/output.txt:214-233; the reference exposes none). */
  public double[] readPrimal(int nvars) {
    double[] x = new double[nvars];
    check(nReadPrimal(handle, nvars, x));
    return x;
  }

  /** (entering, leaving) pairs since the load, oldest first. */
  public int[] pivotLog() {
    long[] cnt = new long[1];
    check(nReadPivotLog(handle, null, 0, cnt));
    int[] pairs = new int[(int) (2 * cnt[0])];
    if (cnt[0] > 0) check(nReadPivotLog(handle, pairs, cnt[0], cnt));
    return pairs;
  }

  // ---- phase-1 support (LPSolver.java:166-246) ---------------------------------------------------------
  /** performDegeneratePivot's scan, LPSolver.java:185-191: first j with |A[row][j]| > epsilon, else -1. */
  public int firstNonzeroInRow(int row) {
    int[] j = new int[1];
    check(nFirstNonzeroInRow(handle, row, j));
    return j[0];
  }

  /** restoreInitialLP, LPSolver.java:205-211 + :235-244: remove column j; positions behind it shift down. */
  public void dropColumn(int j) {
    check(nDropColumn(handle, j));
    n -= 1; // variable ids (and names0) stay; the dropped variable simply no longer has a position
    refreshNames();
  }

  /** restoreInitialLP, LPSolver.java:213-233: c <- 0, v <- 0, then the ops in order.
   *  kind 0: basic variable at row `index`: v += b[index]*coef ; c[j] += (-A[index][j])*coef
   *  kind 1: non-basic variable at column `index`: c[index] += coef */
  public void rebuildObjective(int[] kinds, int[] indexes, double[] coefs) {
    check(nRebuildObjective(handle, kinds, indexes, coefs, kinds.length));
  }

  public boolean hasVariablesNames() {
    return variables != null && coefficients != null;
  }

  @Override
  public void close() {
    if (handle != 0) {
      nDestroy(handle);
      handle = 0;
    }
  }

  // ---- helpers ---------------------------------------------------------------------------------------
  private void check(int rc) {
    if (rc == 0) return;
    String msg = nLastError(handle);
    if (rc == -1) throw new IllegalArgumentException(msg); // Validate.isTrue, LPState.java:288
    throw new IllegalStateException("liblps_b200: " + msg + " (status " + rc + ")");
  }

  private void rememberNames() {
    if (!hasVariablesNames()) return;
    names0 = new String[m + n];
    for (int pos = 0; pos < m + n; pos++) names0[pos] = variables.get(pos);
  }

  /** exchangeIndexes (LPState.java:311-320) happens on the device; mirror it into the two maps. */
  private void refreshNames() {
    if (!hasVariablesNames() || names0 == null) return;
    int[] pos2var = positions();
    variables.clear();
    coefficients.clear();
    for (int pos = 0; pos < pos2var.length; pos++) {
      String name = names0[pos2var[pos]];
      variables.put(pos, name);
      coefficients.put(name, pos);
    }
  }

  static double[] flatten(BigDecimal[][] A, int m, int n) {
    double[] a = new double[Math.max(1, m * n)];
    for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) a[i * n + j] = A[i][j].doubleValue();
    return a;
  }

  static double[] toDoubles(BigDecimal[] x, int len) {
    double[] d = new double[len];
    for (int i = 0; i < len; i++) d[i] = x[i].doubleValue();
    return d;
  }

  // ---- native methods: integration/java/jni/lps_b200_jni.c ------------------------------------------------
  private static native long nCreate(double epsilon, double inf, int device);
  private static native void nDestroy(long h);
  private static native String nLastError(long h);
  private static native int nLoad(long h, int m, int n, double[] a, long lda, double[] b, double[] c, double v);
  private static native int nLoadAux(long h, int m, int n, double[] a, long lda, double[] b);
  private static native int nGetEntering(long h, int[] e);
  private static native int nGetLeaving(long h, int entering, int[] l);
  private static native int nPivot(long h, int entering, int leaving);
  private static native int nRun(long h, long maxPivots, long[] out5, double[] out2);
  private static native int nReadV(long h, double[] v1);
  private static native int nReadB(long h, double[] b);
  private static native int nReadC(long h, double[] c);
  private static native int nReadRow(long h, int i, double[] row);
  private static native int nReadCol(long h, int j, double[] col);
  private static native int nReadPositions(long h, int[] pos2var);
  private static native int nPositionOf(long h, int var, int[] pos1);
  private static native int nReadPrimal(long h, int nvars, double[] x);
  private static native int nReadPivotLog(long h, int[] pairs, long capPairs, long[] count1);
  private static native int nFirstNonzeroInRow(long h, int row, int[] j1);
  private static native int nDropColumn(long h, int j);
  private static native int nRebuildObjective(long h, int[] kinds, int[] indexes, double[] coefs, int nops);
}
