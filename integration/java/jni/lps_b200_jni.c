/* lps_b200_jni.c — JNI glue between lpsolver.LPStateNative (integration/java/lpsolver/LPStateNative.java)
 * and liblps_b200.so (include/lps_b200.h).  One function per native method of the Java class.
 *
 * The build image has no JDK, so the Java side is uncompiled.  THIS file is compiled and executed by the
 * repository's tests against a stand-in <jni.h> (tests/jni_mock/jni.h) whose JNIEnv implements the handful of
 * functions used here; tests/jni_mock/jni_call_order.c drives it in the order LPSolver would.
 *
 * Conventions: primitive arrays are borrowed with GetPrimitiveArrayCritical (no copy; every lps_* call made
 * while a buffer is borrowed is synchronous, so the buffer never outlives the call) and released with
 * JNI_ABORT when the native side only read them, 0 when it wrote them.
 *
 *   build:  gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *               integration/java/jni/lps_b200_jni.c -Llinear_programming_solver_b200 -llps_b200 -o liblps_b200_jni.so
 */
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "lps_b200.h"

#define H(h) ((lps_handle)(intptr_t)(h))
#define FN(name) Java_lpsolver_LPStateNative_##name

static void *borrow(JNIEnv *env, jarray a) { return a ? (*env)->GetPrimitiveArrayCritical(env, a, 0) : 0; }
static void give_back(JNIEnv *env, jarray a, void *p, jint mode) {
  if (a && p) (*env)->ReleasePrimitiveArrayCritical(env, a, p, mode);
}

/* lps_create — new LPState(...), LPState.java:37-112 */
JNIEXPORT jlong JNICALL FN(nCreate)(JNIEnv *env, jclass cls, jdouble epsilon, jdouble inf, jint device) {
  (void)env; (void)cls;
  lps_options o;
  lps_default_options(&o);
  o.epsilon = epsilon;   /* LPState.DEF_EPSILON, LPState.java:20 */
  o.inf = inf;           /* LPState.DEF_INF,     LPState.java:21 */
  o.device = device;
  lps_handle h = 0;
  if (lps_create(&h, &o) != LPS_OK) return 0;
  return (jlong)(intptr_t)h;
}

JNIEXPORT void JNICALL FN(nDestroy)(JNIEnv *env, jclass cls, jlong h) {
  (void)env; (void)cls;
  lps_destroy(H(h));
}

JNIEXPORT jstring JNICALL FN(nLastError)(JNIEnv *env, jclass cls, jlong h) {
  (void)cls;
  return (*env)->NewStringUTF(env, lps_last_error(H(h)));
}

/* lps_load — the constructor's arrays, copied to HBM (the reference aliases them: LPSolver.java:267,270) */
JNIEXPORT jint JNICALL FN(nLoad)(JNIEnv *env, jclass cls, jlong h, jint m, jint n, jdoubleArray a, jlong lda,
                                 jdoubleArray b, jdoubleArray c, jdouble v) {
  (void)cls;
  double *pa = borrow(env, a), *pb = borrow(env, b), *pc = borrow(env, c);
  int rc = lps_load(H(h), m, n, pa, lda, pb, pc, v);
  give_back(env, c, pc, JNI_ABORT);
  give_back(env, b, pb, JNI_ABORT);
  give_back(env, a, pa, JNI_ABORT);
  return rc;
}

/* lps_load_aux — LPSolver.convertIntoAuxLP, LPSolver.java:283-321 */
JNIEXPORT jint JNICALL FN(nLoadAux)(JNIEnv *env, jclass cls, jlong h, jint m, jint n, jdoubleArray a, jlong lda,
                                    jdoubleArray b) {
  (void)cls;
  double *pa = borrow(env, a), *pb = borrow(env, b);
  int rc = lps_load_aux(H(h), m, n, pa, lda, pb);
  give_back(env, b, pb, JNI_ABORT);
  give_back(env, a, pa, JNI_ABORT);
  return rc;
}

/* LPState.getEntering, LPState.java:274-285 */
JNIEXPORT jint JNICALL FN(nGetEntering)(JNIEnv *env, jclass cls, jlong h, jintArray e) {
  (void)cls;
  int out = -1;
  int rc = lps_get_entering(H(h), &out);
  jint *p = borrow(env, e);
  if (p) p[0] = out;
  give_back(env, e, p, 0);
  return rc;
}

/* LPState.getLeaving, LPState.java:287-305 */
JNIEXPORT jint JNICALL FN(nGetLeaving)(JNIEnv *env, jclass cls, jlong h, jint entering, jintArray l) {
  (void)cls;
  int out = -1;
  int rc = lps_get_leaving(H(h), entering, &out);
  jint *p = borrow(env, l);
  if (p) p[0] = out;
  give_back(env, l, p, 0);
  return rc;
}

/* LPState.pivot, LPState.java:114-181 */
JNIEXPORT jint JNICALL FN(nPivot)(JNIEnv *env, jclass cls, jlong h, jint entering, jint leaving) {
  (void)env; (void)cls;
  return lps_pivot(H(h), entering, leaving);
}

/* the loops of LPSolver.simplex (LPSolver.java:101-112) and solveAuxLP (:141-161) */
JNIEXPORT jint JNICALL FN(nRun)(JNIEnv *env, jclass cls, jlong h, jlong max_pivots, jlongArray out5,
                                jdoubleArray out2) {
  (void)cls;
  lps_run_result r;
  memset(&r, 0, sizeof r);
  int rc = lps_run(H(h), max_pivots, &r);
  jlong *pl = borrow(env, out5);
  if (pl) {
    pl[0] = r.verdict;
    pl[1] = r.last_entering;
    pl[2] = r.last_leaving;
    pl[3] = r.npivots;
    pl[4] = r.total_pivots;
  }
  give_back(env, out5, pl, 0);
  jdouble *pd = borrow(env, out2);
  if (pd) {
    pd[0] = r.v;
    pd[1] = r.device_ms;
  }
  give_back(env, out2, pd, 0);
  return rc;
}

/* field reads: LPSolver.java:113 (v), :170 (b), :185 (row), :203-245 (A, b, c) */
JNIEXPORT jint JNICALL FN(nReadV)(JNIEnv *env, jclass cls, jlong h, jdoubleArray v1) {
  (void)cls;
  double v = 0.0;
  int rc = lps_read_v(H(h), &v);
  jdouble *p = borrow(env, v1);
  if (p) p[0] = v;
  give_back(env, v1, p, 0);
  return rc;
}

#define READ_VECTOR(NAME, CALL)                                                              \
  JNIEXPORT jint JNICALL FN(NAME)(JNIEnv *env, jclass cls, jlong h, jdoubleArray out) {      \
    (void)cls;                                                                               \
    jdouble *p = borrow(env, out);                                                           \
    int rc = CALL(H(h), p);                                                                  \
    give_back(env, out, p, 0);                                                               \
    return rc;                                                                               \
  }
READ_VECTOR(nReadB, lps_read_b)
READ_VECTOR(nReadC, lps_read_c)

JNIEXPORT jint JNICALL FN(nReadRow)(JNIEnv *env, jclass cls, jlong h, jint i, jdoubleArray row) {
  (void)cls;
  jdouble *p = borrow(env, row);
  int rc = lps_read_row(H(h), i, p);
  give_back(env, row, p, 0);
  return rc;
}

JNIEXPORT jint JNICALL FN(nReadCol)(JNIEnv *env, jclass cls, jlong h, jint j, jdoubleArray col) {
  (void)cls;
  jdouble *p = borrow(env, col);
  int rc = lps_read_col(H(h), j, p);
  give_back(env, col, p, 0);
  return rc;
}

/* variables / coefficients as a permutation — LPState.java:27-28, :311-320 */
JNIEXPORT jint JNICALL FN(nReadPositions)(JNIEnv *env, jclass cls, jlong h, jintArray pos2var) {
  (void)cls;
  jint *p = borrow(env, pos2var);
  int rc = lps_read_positions(H(h), (int *)p);
  give_back(env, pos2var, p, 0);
  return rc;
}

/* x0 tracking — LPSolver.java:151-155 */
JNIEXPORT jint JNICALL FN(nPositionOf)(JNIEnv *env, jclass cls, jlong h, jint var, jintArray pos1) {
  (void)cls;
  int pos = -1;
  int rc = lps_position_of(H(h), var, &pos);
  jint *p = borrow(env, pos1);
  if (p) p[0] = pos;
  give_back(env, pos1, p, 0);
  return rc;
}

JNIEXPORT jint JNICALL FN(nReadPrimal)(JNIEnv *env, jclass cls, jlong h, jint nvars, jdoubleArray x) {
  (void)cls;
  jdouble *p = borrow(env, x);
  int rc = lps_read_primal(H(h), nvars, p);
  give_back(env, x, p, 0);
  return rc;
}

JNIEXPORT jint JNICALL FN(nReadPivotLog)(JNIEnv *env, jclass cls, jlong h, jintArray pairs, jlong cap,
                                         jlongArray count1) {
  (void)cls;
  int64_t cnt = 0;
  jint *p = borrow(env, pairs);
  int rc = lps_read_pivot_log(H(h), (int *)p, p ? cap : 0, &cnt);
  give_back(env, pairs, p, 0);
  jlong *pc = borrow(env, count1);
  if (pc) pc[0] = cnt;
  give_back(env, count1, pc, 0);
  return rc;
}

/* performDegeneratePivot's scan — LPSolver.java:185-191 */
JNIEXPORT jint JNICALL FN(nFirstNonzeroInRow)(JNIEnv *env, jclass cls, jlong h, jint row, jintArray j1) {
  (void)cls;
  int j = -1;
  int rc = lps_first_nonzero_in_row(H(h), row, &j);
  jint *p = borrow(env, j1);
  if (p) p[0] = j;
  give_back(env, j1, p, 0);
  return rc;
}

/* restoreInitialLP — LPSolver.java:205-211, :235-244 */
JNIEXPORT jint JNICALL FN(nDropColumn)(JNIEnv *env, jclass cls, jlong h, jint j) {
  (void)env; (void)cls;
  return lps_drop_column(H(h), j);
}

/* restoreInitialLP — LPSolver.java:213-233 */
JNIEXPORT jint JNICALL FN(nRebuildObjective)(JNIEnv *env, jclass cls, jlong h, jintArray kinds, jintArray indexes,
                                             jdoubleArray coefs, jint nops) {
  (void)cls;
  if (nops < 0) return LPS_ERR_INVALID;
  enum { kChunk = 256 };
  lps_objective_op ops[kChunk];
  /* the op list is applied in order and lps_rebuild_objective starts from c = 0, v = 0: send it in one call
   * when it fits the stack buffer, else through a heap copy made of the borrowed arrays */
  jint *pk = borrow(env, kinds), *pi = borrow(env, indexes);
  jdouble *pc = borrow(env, coefs);
  int rc = LPS_OK;
  if (nops <= kChunk) {
    for (int k = 0; k < nops; k++) {
      ops[k].kind = pk[k];
      ops[k].index = pi[k];
      ops[k].coef = pc[k];
    }
    rc = lps_rebuild_objective(H(h), ops, nops);
  } else {
    lps_objective_op *big = (lps_objective_op *)malloc((size_t)nops * sizeof *big);
    if (!big) rc = LPS_ERR_NOMEM;
    else {
      for (int k = 0; k < nops; k++) {
        big[k].kind = pk[k];
        big[k].index = pi[k];
        big[k].coef = pc[k];
      }
      rc = lps_rebuild_objective(H(h), big, nops);
      free(big);
    }
  }
  give_back(env, coefs, pc, JNI_ABORT);
  give_back(env, indexes, pi, JNI_ABORT);
  give_back(env, kinds, pk, JNI_ABORT);
  return rc;
}
