/* Drives integration/java/jni/lps_b200_jni.c exactly as lpsolver.LPStateNative + the patched LPSolver would
 * (integration/java/LPSolver.patch): TEST INFRASTRUCTURE, needs a GPU (the glue calls liblps_b200.so).
 *
 *   phase 2 only:   nCreate -> nLoad -> nRun(-1) -> nReadV / nReadB / nReadC / nReadPositions / nReadPrimal
 *   with phase 1:   nCreate -> nLoadAux -> nPivot(x0, minInB) -> nRun(-1) -> nPositionOf(x0) -> nReadB
 *                   [-> nFirstNonzeroInRow -> nPivot] -> nDropColumn -> nRebuildObjective -> nRun(-1) -> nReadV
 *   explicit loop:  nGetEntering / nGetLeaving / nPivot until -1 (the unpatched loop of LPSolver.java:101-112)
 *
 * Input: a tiny text file "m n maximize" + A rows + b + c.  Output: one line of JSON with the verdict, v, the
 * pivot log, b, c and the positions, compared by tests/test_jni_glue.py with the binary64 oracle.  Every borrowed
 * array must have been released by the time a native function returns. */
#include <jni.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define FN(name) Java_lpsolver_LPStateNative_##name
jlong FN(nCreate)(JNIEnv *, jclass, jdouble, jdouble, jint);
void FN(nDestroy)(JNIEnv *, jclass, jlong);
jstring FN(nLastError)(JNIEnv *, jclass, jlong);
jint FN(nLoad)(JNIEnv *, jclass, jlong, jint, jint, jdoubleArray, jlong, jdoubleArray, jdoubleArray, jdouble);
jint FN(nLoadAux)(JNIEnv *, jclass, jlong, jint, jint, jdoubleArray, jlong, jdoubleArray);
jint FN(nGetEntering)(JNIEnv *, jclass, jlong, jintArray);
jint FN(nGetLeaving)(JNIEnv *, jclass, jlong, jint, jintArray);
jint FN(nPivot)(JNIEnv *, jclass, jlong, jint, jint);
jint FN(nRun)(JNIEnv *, jclass, jlong, jlong, jlongArray, jdoubleArray);
jint FN(nReadV)(JNIEnv *, jclass, jlong, jdoubleArray);
jint FN(nReadB)(JNIEnv *, jclass, jlong, jdoubleArray);
jint FN(nReadC)(JNIEnv *, jclass, jlong, jdoubleArray);
jint FN(nReadRow)(JNIEnv *, jclass, jlong, jint, jdoubleArray);
jint FN(nReadPositions)(JNIEnv *, jclass, jlong, jintArray);
jint FN(nPositionOf)(JNIEnv *, jclass, jlong, jint, jintArray);
jint FN(nReadPrimal)(JNIEnv *, jclass, jlong, jint, jdoubleArray);
jint FN(nReadPivotLog)(JNIEnv *, jclass, jlong, jintArray, jlong, jlongArray);
jint FN(nFirstNonzeroInRow)(JNIEnv *, jclass, jlong, jint, jintArray);
jint FN(nDropColumn)(JNIEnv *, jclass, jlong, jint);
jint FN(nRebuildObjective)(JNIEnv *, jclass, jlong, jintArray, jintArray, jdoubleArray, jint);

/* ---- the mock JVM ------------------------------------------------------------------------------------ */
static int outstanding = 0;
static void *get_critical(JNIEnv *env, jarray a, jboolean *is_copy) {
  (void)env;
  if (is_copy) *is_copy = 0;
  a->borrowed++;
  outstanding++;
  return a->data;
}
static void release_critical(JNIEnv *env, jarray a, void *p, jint mode) {
  (void)env;
  if (p != a->data || a->borrowed <= 0) { fprintf(stderr, "mock JVM: bad ReleasePrimitiveArrayCritical\n"); exit(3); }
  a->borrowed--;
  outstanding--;
  if (mode == 0) a->writes_committed++;
}
static jstring new_string(JNIEnv *env, const char *s) { (void)env; return strdup(s ? s : ""); }
static const struct JNINativeInterface_ table = {get_critical, release_critical, new_string};
static JNIEnv env_obj = &table;
static JNIEnv *env = &env_obj;

static jarray arr(jint len, size_t elem) {
  jarray a = calloc(1, sizeof *a);
  a->length = len;
  a->data = calloc(len > 0 ? len : 1, elem);
  return a;
}
#define D(a) ((double *)(a)->data)
#define I(a) ((jint *)(a)->data)
#define L(a) ((jlong *)(a)->data)
static void no_leak(const char *where) {
  if (outstanding != 0) { fprintf(stderr, "mock JVM: %d critical borrow(s) still outstanding after %s\n", outstanding, where); exit(4); }
}
static void ck(jint rc, jlong h, const char *what) {
  no_leak(what);
  if (rc != 0) { fprintf(stderr, "%s failed: %d (%s)\n", what, rc, FN(nLastError)(env, 0, h)); exit(2); }
}
static void print_d(const char *k, jarray a, int n) {
  printf("\"%s\": [", k);
  for (int i = 0; i < n; i++) printf("%s%.17g", i ? ", " : "", D(a)[i]);
  printf("]");
}
static void print_i(const char *k, jarray a, int n) {
  printf("\"%s\": [", k);
  for (int i = 0; i < n; i++) printf("%s%d", i ? ", " : "", I(a)[i]);
  printf("]");
}

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: jni_call_order <lp.txt> run|loop|phase1\n"); return 1; }
  FILE *f = fopen(argv[1], "r");
  if (!f) { perror(argv[1]); return 1; }
  int m, n, maximize;
  if (fscanf(f, "%d %d %d", &m, &n, &maximize) != 3) return 1;
  jarray A = arr(m * n, 8), b = arr(m, 8), c = arr(n, 8);
  for (int k = 0; k < m * n; k++) if (fscanf(f, "%lf", &D(A)[k]) != 1) return 1;
  for (int k = 0; k < m; k++) if (fscanf(f, "%lf", &D(b)[k]) != 1) return 1;
  for (int k = 0; k < n; k++) if (fscanf(f, "%lf", &D(c)[k]) != 1) return 1;
  fclose(f);
  const char *mode = argv[2];
  jarray i1 = arr(1, 4), l5 = arr(5, 8), d2 = arr(2, 8), d1 = arr(1, 8), cnt = arr(1, 8);
  jlong h = FN(nCreate)(env, 0, 1e-9, 1e50, -1);
  if (!h) { fprintf(stderr, "nCreate: no CUDA device\n"); return 5; }
  int nn = n;      /* columns of the tableau the handle currently holds */
  int verdict = -1;
  if (!strcmp(mode, "phase1")) {
    /* LPSolver.initializeSimplex -> convertIntoAuxLP -> solveAuxLP -> handleInitialization -> restoreInitialLP */
    int k = -1;    /* minInB: first index of the strict minimum (LPSolver.java:375-386) */
    double best = 1e50;
    for (int i = 0; i < m; i++) if (D(b)[i] < best) { best = D(b)[i]; k = i; }
    if (k < 0 || D(b)[k] >= 0) { fprintf(stderr, "phase1 mode needs a negative right-hand side\n"); return 1; }
    ck(FN(nLoadAux)(env, 0, h, m, n, A, n, b), h, "nLoadAux");
    nn = n + 1;
    ck(FN(nPivot)(env, 0, h, n, k), h, "nPivot(x0, minInB)");                    /* LPSolver.java:138 */
    ck(FN(nRun)(env, 0, h, -1, l5, d2), h, "nRun(aux)");                         /* :141-161 */
    if (L(l5)[0] == 2) { printf("{\"verdict\": \"Auxiliary lp is unbounded\"}\n"); return 0; }
    ck(FN(nPositionOf)(env, 0, h, n, i1), h, "nPositionOf(x0)");                 /* x0 is variable id n */
    int x0 = I(i1)[0];
    jarray bb = arr(m, 8);
    ck(FN(nReadB)(env, 0, h, bb), h, "nReadB");
    double x0v = (x0 < nn) ? 0.0 : D(bb)[x0 - nn];
    if (fabs(x0v) > 1e-9) { printf("{\"verdict\": \"This linear program is infeasible\"}\n"); return 0; }   /* :169-174 */
    if (x0 >= nn) {                                                              /* performDegeneratePivot, :182-198 */
      ck(FN(nFirstNonzeroInRow)(env, 0, h, x0 - nn, i1), h, "nFirstNonzeroInRow");
      int e = I(i1)[0];
      if (e < 0) { printf("{\"verdict\": \"Can't perform degenerate pivot\"}\n"); return 0; }
      ck(FN(nPivot)(env, 0, h, e, x0 - nn), h, "nPivot(degenerate)");
      x0 = e;
    }
    /* restoreInitialLP (:200-246) with the index fix (fix_restore_index = true): the objective
     * rebuild in ascending variable order, positions taken AFTER the column drop */
    jarray pos = arr(m + nn, 4);
    ck(FN(nReadPositions)(env, 0, h, pos), h, "nReadPositions");
    ck(FN(nDropColumn)(env, 0, h, x0), h, "nDropColumn");
    nn = n;
    jarray kinds = arr(n, 4), idx = arr(n, 4), coef = arr(n, 8);
    for (int var = 0; var < n; var++) {
      int p = -1;
      for (int q = 0; q < m + n + 1; q++) if (I(pos)[q] == var) { p = q; break; }
      if (p > x0) p -= 1;                         /* positions behind the dropped column shift down */
      if (p >= n) { I(kinds)[var] = 0; I(idx)[var] = p - n; }
      else { I(kinds)[var] = 1; I(idx)[var] = p; }
      D(coef)[var] = D(c)[var];
    }
    ck(FN(nRebuildObjective)(env, 0, h, kinds, idx, coef, n), h, "nRebuildObjective");
    ck(FN(nRun)(env, 0, h, -1, l5, d2), h, "nRun(phase 2)");
    verdict = (int)L(l5)[0];
  } else if (!strcmp(mode, "loop")) {
    ck(FN(nLoad)(env, 0, h, m, n, A, n, b, c, 0.0), h, "nLoad");
    verdict = 1;
    for (;;) {                                                                   /* LPSolver.java:101-112 */
      ck(FN(nGetEntering)(env, 0, h, i1), h, "nGetEntering");
      int e = I(i1)[0];
      if (e == -1) break;
      ck(FN(nGetLeaving)(env, 0, h, e, i1), h, "nGetLeaving");
      int l = I(i1)[0];
      if (l == -1) { verdict = 2; break; }
      ck(FN(nPivot)(env, 0, h, e, l), h, "nPivot");
    }
  } else {
    ck(FN(nLoad)(env, 0, h, m, n, A, n, b, c, 0.0), h, "nLoad");
    ck(FN(nRun)(env, 0, h, -1, l5, d2), h, "nRun");
    verdict = (int)L(l5)[0];
  }
  jarray ob = arr(m, 8), oc = arr(nn, 8), op = arr(m + nn, 4), ox = arr(n, 8);
  ck(FN(nReadV)(env, 0, h, d1), h, "nReadV");
  ck(FN(nReadB)(env, 0, h, ob), h, "nReadB");
  ck(FN(nReadC)(env, 0, h, oc), h, "nReadC");
  ck(FN(nReadPositions)(env, 0, h, op), h, "nReadPositions");
  ck(FN(nReadPrimal)(env, 0, h, n, ox), h, "nReadPrimal");
  ck(FN(nReadPivotLog)(env, 0, h, 0, 0, cnt), h, "nReadPivotLog(count)");
  jlong np = L(cnt)[0];
  jarray lg = arr((jint)(2 * np), 4);
  if (np > 0) ck(FN(nReadPivotLog)(env, 0, h, lg, np, cnt), h, "nReadPivotLog");
  printf("{\"verdict\": %d, \"v\": %.17g, \"pivots\": %lld, ", verdict, D(d1)[0], (long long)np);
  print_i("log", lg, (int)(2 * np)); printf(", ");
  print_d("b", ob, m); printf(", ");
  print_d("c", oc, nn); printf(", ");
  print_i("positions", op, m + nn); printf(", ");
  print_d("primal", ox, n);
  printf("}\n");
  FN(nDestroy)(env, 0, h);
  return 0;
}
