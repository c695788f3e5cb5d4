/*
 * Tier F — binary64 CPU twin of the reference's pivot path.   TEST INFRASTRUCTURE ONLY.
 *
 * This file belongs to the oracle (see oracle/__init__.py): it is the checker for the CUDA
 * kernels and the timed CPU baseline of bench.py.  Nothing in the product links or calls it.
 *
 * It restates, in IEEE binary64 with one rounding per operation (compile with
 * -ffp-contract=off: the multiply and the subtract are rounded separately, as the
 * reference rounds `multiply(…, rounder)` and `subtract(…, rounder)` separately), the
 * following pieces of Toptachamann/Linear_Programming_Solver, src/main/java/lpsolver/:
 *
 *   tf_get_entering   LPState.java:274-285   first index with c[i] > eps
 *   tf_get_leaving    LPState.java:287-305   lowest-index min of b[i]/A[i][e], A[i][e] >= eps,
 *                                            ratio < INF
 *   tf_pivot          LPState.java:133-181   pivotSequentially; with nthreads > 1 the three
 *                     LPState.java:184-272   barrier-separated phases of pivotConcurrently,
 *                                            each split into nthreads contiguous blocks
 *   tf_run            LPSolver.java:101-112  the simplex loop (and the loop part of
 *                     LPSolver.java:141-161  solveAuxLP; x0 is tracked through pos2var)
 *   tf_min_in_b       LPSolver.java:375-386
 *
 * Parity: pinned, through tests/test_oracle_golden.py and tests/test_tier_f.py, to the
 * reference's Spock vectors (all exactly representable in binary64) and to the Python
 * restatement in simplex_ref.py, which in turn is checked in decimal-15 arithmetic.
 * The reference itself is Java and cannot be built here (no JVM): there is no oracle/_ref.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TF_OPTIMAL 0
#define TF_UNBOUNDED 1
#define TF_PIVOT_CAP 2

int tf_version(void) { return 1; }

int tf_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int tf_get_entering(const double *c, int n, double eps) {
  for (int i = 0; i < n; i++)
    if (c[i] > eps) return i;
  return -1;
}

int tf_get_leaving(const double *A, long lda, const double *b, int m, int e, double eps,
                   double inf) {
  int leaving = -1;
  double min_slack = inf;
  for (int i = 0; i < m; i++) {
    double aie = A[(long)i * lda + e];
    double slack = (aie < eps) ? inf : b[i] / aie;
    if (slack < min_slack) {
      min_slack = slack;
      leaving = i;
    }
  }
  return leaving;
}

int tf_min_in_b(const double *b, int m, double inf) {
  double cur = inf;
  int idx = -1;
  for (int i = 0; i < m; i++)
    if (cur > b[i]) {
      cur = b[i];
      idx = i;
    }
  return idx;
}

/* row[j] <- row[j] - a * prow[j] for j in [lo, hi): separate mul and sub roundings. */
__attribute__((target_clones("avx512f", "avx2", "default"))) static void
axmy(double *restrict row, const double *restrict prow, double a, long lo, long hi) {
  for (long j = lo; j < hi; j++) row[j] = row[j] - a * prow[j];
}

__attribute__((target_clones("avx512f", "avx2", "default"))) static void
divrow(double *restrict row, double p, long lo, long hi) {
  for (long j = lo; j < hi; j++) row[j] = row[j] / p;
}

void tf_pivot(double *A, long lda, double *b, double *c, double *v, int m, int n, int e, int l,
              int nthreads) {
  double *prow = A + (long)l * lda;
  const double p = prow[e];
  if (nthreads < 1) nthreads = 1;
  /* phase 1: pivot row (LPState.java:139-146 / :194-213) */
  prow[e] = 1.0 / p;
#pragma omp parallel for schedule(static) num_threads(nthreads) if (nthreads > 1)
  for (int k = 0; k < nthreads; k++) {
    long from = ((long)k * n) / nthreads, to = ((long)(k + 1) * n) / nthreads;
    if (e >= from && e < to) {
      divrow(prow, p, from, e);
      divrow(prow, p, e + 1, to);
    } else {
      divrow(prow, p, from, to);
    }
  }
  b[l] = b[l] / p;
  const double b_ent = b[l];
  /* phase 2: other rows (LPState.java:151-166 / :218-245) */
#pragma omp parallel for schedule(static) num_threads(nthreads) if (nthreads > 1)
  for (int k = 0; k < nthreads; k++) {
    long from = ((long)k * m) / nthreads, to = ((long)(k + 1) * m) / nthreads;
    for (long i = from; i < to; i++) {
      if (i == l) continue;
      double *row = A + i * lda;
      const double a = row[e];
      axmy(row, prow, a, 0, e);
      axmy(row, prow, a, e + 1, n);
      row[e] = -(a / p);
      b[i] = b[i] - a * b_ent;
    }
  }
  /* phase 3: objective (LPState.java:170-178 / :249-264) */
  const double ce = c[e];
  *v = *v + b[l] * ce;
  c[e] = -(ce / p);
#pragma omp parallel for schedule(static) num_threads(nthreads) if (nthreads > 1)
  for (int k = 0; k < nthreads; k++) {
    long from = ((long)k * n) / nthreads, to = ((long)(k + 1) * n) / nthreads;
    if (e >= from && e < to) {
      axmy(c, prow, ce, from, e);
      axmy(c, prow, ce, e + 1, to);
    } else {
      axmy(c, prow, ce, from, to);
    }
  }
}

/* exchangeIndexes (LPState.java:311-320) on an index permutation instead of name maps:
 * pos2var[pos] = id of the variable sitting at position pos (0..n-1 non-basic, n..n+m-1 basic). */
static void swap_positions(int *pos2var, int n, int e, int l) {
  if (!pos2var) return;
  int t = pos2var[e];
  pos2var[e] = pos2var[n + l];
  pos2var[n + l] = t;
}

void tf_pivot_tracked(double *A, long lda, double *b, double *c, double *v, int m, int n, int e,
                      int l, int nthreads, int *pos2var) {
  tf_pivot(A, lda, b, c, v, m, n, e, l, nthreads);
  swap_positions(pos2var, n, e, l);
}

/* The loop of LPSolver.simplex (LPSolver.java:101-112).  log holds (e,l) pairs, log_cap pairs
 * at most (pivots beyond the cap are still executed, just not recorded). */
int tf_run(double *A, long lda, double *b, double *c, double *v, int m, int n, double eps,
           double inf, long max_pivots, int *log, long log_cap, long *npivots, int *pos2var,
           int nthreads) {
  long k = 0;
  int status = TF_OPTIMAL;
  for (;;) {
    int e = tf_get_entering(c, n, eps);
    if (e == -1) break;
    int l = tf_get_leaving(A, lda, b, m, e, eps, inf);
    if (l == -1) {
      status = TF_UNBOUNDED;
      break;
    }
    if (max_pivots >= 0 && k >= max_pivots) {
      status = TF_PIVOT_CAP;
      break;
    }
    tf_pivot(A, lda, b, c, v, m, n, e, l, nthreads);
    swap_positions(pos2var, n, e, l);
    if (log && k < log_cap) {
      log[2 * k] = e;
      log[2 * k + 1] = l;
    }
    k++;
  }
  if (npivots) *npivots = k;
  return status;
}

/* ---- synthetic inputs: counter-based, so any shard can regenerate any cell.
 * u(seed,k) = ((splitmix64(seed ^ (k * GOLDEN)) >> 44) + 1) / 2^20  in (0,1], dyadic. */
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

double tf_u(uint64_t seed, uint64_t k) {
  uint64_t h = splitmix64(seed ^ (k * 0x9E3779B97F4A7C15ULL));
  return (double)((h >> 44) + 1) * (1.0 / 1048576.0);
}

/* A[i][j] = u(i*n+j) for rows [row0,row1) of an m x n matrix, written at A (row row0 first). */
void tf_fill_u(double *A, long lda, long row0, long row1, long n, uint64_t seed, uint64_t base,
               int nthreads) {
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(static) num_threads(nthreads) if (nthreads > 1)
  for (long i = row0; i < row1; i++) {
    double *row = A + (i - row0) * lda;
    for (long j = 0; j < n; j++) row[j] = tf_u(seed, base + (uint64_t)i * (uint64_t)n + (uint64_t)j);
  }
}
