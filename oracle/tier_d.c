/*
 * Tier D in C — the reference's own arithmetic, fast enough for BASELINE-sized checks.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for the CUDA path, never linked
 * or called by the product.
 *
 * Restates java.math.BigDecimal with MathContext(15, RoundingMode.HALF_UP) as used on the hot
 * path of Toptachamann/Linear_Programming_Solver (LPState.java:18 and every
 * divide/multiply/subtract/add(…, rounder) call at LPState.java:139-177,297): each operation
 * computes the EXACT result and rounds it once to 15 significant decimal digits, half up;
 * compareTo (LPState.java:278,294,299) is exact.  Inputs are not rounded on entry.  On top of that
 * arithmetic: getEntering (LPState.java:274-285), getLeaving (:287-305), pivotSequentially
 * (:133-181) and the loop of LPSolver.simplex (LPSolver.java:101-112).
 *
 * A number is sign * coef * 10^exp with coef < 10^38 held in an unsigned __int128; exact
 * intermediates use a 256-bit integer.  Parity pinning: every operation is cross-checked against
 * Python's `decimal` (the same General Decimal Arithmetic semantics) on random and adversarial
 * operands, and whole solves against oracle/simplex_ref.py with Dec15 — tests/test_tier_d.py.
 * Limits (asserted): operands of add/sub/div have at most 30 significant digits, which covers
 * binary64 inputs with up to 27 fractional bits (the synthetic generator uses 20).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef struct { uint64_t w[4]; } u256; /* little-endian limbs */

typedef struct {
  u128 coef;
  int32_t exp;
  int32_t neg; /* 1 = negative; zero is always non-negative (BigDecimal has no signed zero) */
} dec;

static u128 P10_128[39];
static u256 P10_256[78];
static int g_init = 0;

/* ---- u256 helpers ---- */
static u256 u256_from128(u128 x) { u256 r = {{(uint64_t)x, (uint64_t)(x >> 64), 0, 0}}; return r; }
static int u256_is_zero(const u256 *a) { return !(a->w[0] | a->w[1] | a->w[2] | a->w[3]); }
static int u256_cmp(const u256 *a, const u256 *b) {
  for (int i = 3; i >= 0; i--) {
    if (a->w[i] != b->w[i]) return a->w[i] < b->w[i] ? -1 : 1;
  }
  return 0;
}
static u256 u256_add(const u256 *a, const u256 *b) {
  u256 r; u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a->w[i] + b->w[i]; r.w[i] = (uint64_t)c; c >>= 64; }
  return r;
}
static u256 u256_sub(const u256 *a, const u256 *b) { /* a >= b */
  u256 r; u128 borrow = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a->w[i] - b->w[i] - borrow;
    r.w[i] = (uint64_t)d;
    borrow = (d >> 64) & 1;
  }
  return r;
}
static u256 u256_mul128(u128 a, u128 b) {
  uint64_t al = (uint64_t)a, ah = (uint64_t)(a >> 64), bl = (uint64_t)b, bh = (uint64_t)(b >> 64);
  u128 p0 = (u128)al * bl, p1 = (u128)al * bh, p2 = (u128)ah * bl, p3 = (u128)ah * bh;
  u256 r;
  r.w[0] = (uint64_t)p0;
  u128 mid = (p0 >> 64) + (uint64_t)p1 + (uint64_t)p2;
  r.w[1] = (uint64_t)mid;
  u128 hi = (mid >> 64) + (p1 >> 64) + (p2 >> 64) + (uint64_t)p3;
  r.w[2] = (uint64_t)hi;
  r.w[3] = (uint64_t)((hi >> 64) + (p3 >> 64));
  return r;
}
static u256 u256_mul_small(const u256 *a, uint64_t m) {
  u256 r; u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a->w[i] * m; r.w[i] = (uint64_t)c; c >>= 64; }
  return r;
}
static uint64_t u256_divmod_small(u256 *a, uint64_t d) { /* a /= d, returns remainder */
  u128 rem = 0;
  for (int i = 3; i >= 0; i--) {
    u128 cur = (rem << 64) | a->w[i];
    a->w[i] = (uint64_t)(cur / d);
    rem = cur % d;
  }
  return (uint64_t)rem;
}
static int u256_fits128(const u256 *a) { return !(a->w[2] | a->w[3]); }
static u128 u256_to128(const u256 *a) { return ((u128)a->w[1] << 64) | a->w[0]; }

static void td_init(void) {
  if (g_init) return;
  P10_128[0] = 1;
  for (int i = 1; i < 39; i++) P10_128[i] = P10_128[i - 1] * 10;
  P10_256[0] = u256_from128(1);
  for (int i = 1; i < 78; i++) P10_256[i] = u256_mul_small(&P10_256[i - 1], 10);
  g_init = 1;
}

static int digits128(u128 x) { /* 0 -> 0 digits */
  int lo = 0, hi = 39;         /* smallest d with x < 10^d */
  if (x == 0) return 0;
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (mid < 39 && x < P10_128[mid]) hi = mid; else lo = mid + 1;
  }
  return lo;
}
static int digits256(const u256 *x) {
  if (u256_fits128(x)) {
    u128 v = u256_to128(x);
    if (v < P10_128[38]) return digits128(v);
  }
  int lo = 38, hi = 78;
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (mid < 78 && u256_cmp(x, &P10_256[mid]) < 0) hi = mid; else lo = mid + 1;
  }
  return lo;
}

/* divide by 10^k (k >= 0), truncating */
static void u256_shr10(u256 *a, int k) {
  while (k > 0) {
    int s = k > 19 ? 19 : k;
    u256_divmod_small(a, (uint64_t)P10_128[s]);
    k -= s;
  }
}

/* exact value (S, exp) -> rounded to 15 significant digits, HALF_UP: only the first discarded
 * digit matters (>= 5 rounds away from zero). */
static dec round15_256(u256 S, int32_t exp, int neg) {
  dec r;
  if (u256_is_zero(&S)) { r.coef = 0; r.exp = exp; r.neg = 0; return r; }
  int d = digits256(&S);
  if (d > 15) {
    u256_shr10(&S, d - 16);                 /* 16 digits left */
    uint64_t t = S.w[0];
    uint64_t last = t % 10, q = t / 10;
    if (last >= 5) q++;
    exp += d - 15;
    if (q == 1000000000000000ULL) { q = 100000000000000ULL; exp++; }
    r.coef = q;
  } else {
    r.coef = u256_to128(&S);
  }
  r.exp = exp;
  r.neg = neg;
  return r;
}
static dec round15_128(u128 S, int32_t exp, int neg) {
  dec r;
  if (S == 0) { r.coef = 0; r.exp = exp; r.neg = 0; return r; }
  int d = digits128(S);
  if (d > 15) {
    u128 t = S / P10_128[d - 16];
    uint64_t last = (uint64_t)(t % 10), q = (uint64_t)(t / 10);
    if (last >= 5) q++;
    exp += d - 15;
    if (q == 1000000000000000ULL) { q = 100000000000000ULL; exp++; }
    r.coef = q;
  } else {
    r.coef = S;
  }
  r.exp = exp;
  r.neg = neg;
  return r;
}

static dec dec_neg(dec a) { if (a.coef) a.neg = !a.neg; return a; }

static dec dec_mul(dec a, dec b) {
  int neg = a.neg ^ b.neg;
  if (a.coef == 0 || b.coef == 0) { dec z = {0, a.exp + b.exp, 0}; return z; }
  if (a.coef < P10_128[19] && b.coef < P10_128[19]) return round15_128(a.coef * b.coef, a.exp + b.exp, neg);
  return round15_256(u256_mul128(a.coef, b.coef), a.exp + b.exp, neg);
}

/* a + b (exact, then rounded) */
static dec dec_add(dec a, dec b) {
  if (a.coef == 0) return round15_128(b.coef, b.exp, b.neg);
  if (b.coef == 0) return round15_128(a.coef, a.exp, a.neg);
  if (a.exp < b.exp) { dec t = a; a = b; b = t; }     /* a has the larger exponent */
  int shift = a.exp - b.exp;
  int da = digits128(a.coef);
  if (shift <= 38 - da) {                              /* everything fits 128 bits */
    u128 X = a.coef * P10_128[shift], Y = b.coef;
    if (a.neg == b.neg) {
      u128 S = X + Y;
      if (S >= X) return round15_128(S, b.exp, a.neg);  /* no wrap */
    } else {
      if (X >= Y) return round15_128(X - Y, b.exp, X == Y ? 0 : a.neg);
      return round15_128(Y - X, b.exp, b.neg);
    }
  }
  u256 X, Y;
  int32_t exp;
  if (da + shift <= 76) {                              /* exact alignment fits 256 bits */
    X = u256_from128(a.coef);
    int s = shift;
    while (s > 0) { int k = s > 19 ? 19 : s; X = u256_mul_small(&X, (uint64_t)P10_128[k]); s -= k; }
    Y = u256_from128(b.coef);
    exp = b.exp;
  } else {
    /* b lies entirely below 10^(a.exp-k): it cannot change any kept digit, it only decides which
     * side of a rounding boundary a sits on.  Give a at least 21 digits (so the result IS rounded)
     * and let b act as one unit in the last place.  (|b| < 10^(b.exp+38) <= 10^(a.exp-k) because
     * shift > 76 - da >= 38 + k for da <= 30.) */
    int k = da >= 16 ? 5 : 21 - da;
    X = u256_from128(a.coef);
    int s = k;
    while (s > 0) { int q = s > 19 ? 19 : s; X = u256_mul_small(&X, (uint64_t)P10_128[q]); s -= q; }
    Y = u256_from128(1);
    exp = a.exp - k;
  }
  if (a.neg == b.neg) return round15_256(u256_add(&X, &Y), exp, a.neg);
  int c = u256_cmp(&X, &Y);
  if (c == 0) { dec z = {0, exp, 0}; return z; }
  if (c > 0) return round15_256(u256_sub(&X, &Y), exp, a.neg);
  return round15_256(u256_sub(&Y, &X), exp, b.neg);
}
static dec dec_sub(dec a, dec b) { return dec_add(a, dec_neg(b)); }

static dec dec_div(dec a, dec b) { /* b != 0 */
  int neg = a.neg ^ b.neg;
  if (a.coef == 0) { dec z = {0, a.exp - b.exp, 0}; return z; }
  u128 Q = a.coef / b.coef, r = a.coef % b.coef;
  int32_t e = a.exp - b.exp;
  while (Q < P10_128[16] && r != 0) {    /* long division until >= 17 digits or exact */
    r *= 10;
    Q = Q * 10 + r / b.coef;
    r %= b.coef;
    e--;
  }
  return round15_128(Q, e, neg);
}

static int dec_cmp(dec a, dec b) {
  if (a.coef == 0 && b.coef == 0) return 0;
  if (a.coef == 0) return b.neg ? 1 : -1;
  if (b.coef == 0) return a.neg ? -1 : 1;
  if (a.neg != b.neg) return a.neg ? -1 : 1;
  int sign = a.neg ? -1 : 1;
  int aa = a.exp + digits128(a.coef), ab = b.exp + digits128(b.coef);   /* adjusted exponent + 1 */
  if (aa != ab) return aa < ab ? -sign : sign;
  /* same magnitude class: align exactly (shift < 39) */
  u256 X = u256_from128(a.coef), Y = u256_from128(b.coef);
  int s = a.exp - b.exp;
  while (s > 0) { int k = s > 19 ? 19 : s; X = u256_mul_small(&X, (uint64_t)P10_128[k]); s -= k; }
  while (s < 0) { int k = -s > 19 ? 19 : -s; Y = u256_mul_small(&Y, (uint64_t)P10_128[k]); s += k; }
  int c = u256_cmp(&X, &Y);
  return c * sign;
}

/* exact conversion of a binary64 (needs <= 38 decimal digits: dyadic with few fractional bits) */
static int dec_from_double(double x, dec *out) {
  dec r = {0, 0, 0};
  if (x == 0.0) { *out = r; return 0; }
  if (!isfinite(x)) return -1;
  r.neg = x < 0;
  int e2;
  double m = frexp(fabs(x), &e2);               /* |x| = m * 2^e2, 0.5 <= m < 1 */
  uint64_t mant = (uint64_t)ldexp(m, 53);
  e2 -= 53;
  while ((mant & 1) == 0) { mant >>= 1; e2++; }
  u128 c = mant;
  if (e2 >= 0) {
    if (e2 > 60) return -1;
    c <<= e2;
    if (c >= P10_128[38]) return -1;
  } else {
    int k = -e2;                                /* x = mant * 5^k / 10^k */
    for (int i = 0; i < k; i++) {
      if (c >= P10_128[37]) return -1;
      c *= 5;
    }
    r.exp = -k;
  }
  r.coef = c;
  *out = r;
  return 0;
}

static void dec_to_string(dec a, char *buf, size_t cap) {
  char digs[48];
  int n = 0;
  u128 c = a.coef;
  if (c == 0) { digs[n++] = '0'; }
  while (c) { digs[n++] = (char)('0' + (int)(c % 10)); c /= 10; }
  size_t p = 0;
  if (a.neg && p < cap - 1) buf[p++] = '-';
  for (int i = n - 1; i >= 0 && p < cap - 1; i--) buf[p++] = digs[i];
  snprintf(buf + p, cap - p, "E%d", a.exp);
}
static double dec_to_double(dec a) {
  char buf[80];
  dec_to_string(a, buf, sizeof buf);
  return strtod(buf, NULL);
}

/* ---- exported scalar ops (for the cross-check against Python decimal) ---- */
typedef struct { uint64_t lo, hi; int32_t exp, neg; } dec_io;
static dec in(dec_io x) { dec d = {((u128)x.hi << 64) | x.lo, x.exp, x.neg}; return d; }
static dec_io out(dec d) { dec_io x = {(uint64_t)d.coef, (uint64_t)(d.coef >> 64), d.exp, d.neg}; return x; }
void td_op(int op, const dec_io *a, const dec_io *b, dec_io *r) {
  td_init();
  dec x = in(*a), y = in(*b), z;
  switch (op) {
    case 0: z = dec_mul(x, y); break;
    case 1: z = dec_add(x, y); break;
    case 2: z = dec_sub(x, y); break;
    case 3: z = dec_div(x, y); break;
    default: z.coef = 0; z.exp = 0; z.neg = 0; break;
  }
  *r = out(z);
}
int td_cmp(const dec_io *a, const dec_io *b) { td_init(); return dec_cmp(in(*a), in(*b)); }
int td_from_double(double x, dec_io *r) { td_init(); dec d; int rc = dec_from_double(x, &d); if (!rc) *r = out(d); return rc; }

/* ---- LPState in decimal-15 ---- */
typedef struct {
  int m, n;
  dec *A, *b, *c;
  dec v, eps, inf;
  int *pos2var;
} td_state;

#define TD_OPTIMAL 0
#define TD_UNBOUNDED 1
#define TD_PIVOT_CAP 2

td_state *td_create(int m, int n, const double *A, long lda, const double *b, const double *c) {
  td_init();
  td_state *s = calloc(1, sizeof *s);
  s->m = m; s->n = n;
  s->A = malloc(sizeof(dec) * (size_t)(m > 0 ? m : 1) * (size_t)(n > 0 ? n : 1));
  s->b = malloc(sizeof(dec) * (size_t)(m > 0 ? m : 1));
  s->c = malloc(sizeof(dec) * (size_t)(n > 0 ? n : 1));
  s->pos2var = malloc(sizeof(int) * (size_t)(m + n + 1));
  int bad = 0;
  for (int i = 0; i < m; i++) {
    for (int j = 0; j < n; j++) bad |= dec_from_double(A[(long)i * lda + j], &s->A[(size_t)i * n + j]);
    bad |= dec_from_double(b[i], &s->b[i]);
  }
  for (int j = 0; j < n; j++) bad |= dec_from_double(c[j], &s->c[j]);
  for (int k = 0; k < m + n; k++) s->pos2var[k] = k;
  s->v.coef = 0; s->v.exp = 0; s->v.neg = 0;
  s->eps.coef = 1; s->eps.exp = -9; s->eps.neg = 0;    /* LPState.DEF_EPSILON, LPState.java:20 */
  s->inf.coef = 1; s->inf.exp = 50; s->inf.neg = 0;    /* LPState.DEF_INF,     LPState.java:21 */
  if (bad) { free(s->A); free(s->b); free(s->c); free(s->pos2var); free(s); return NULL; }
  return s;
}
void td_destroy(td_state *s) { if (s) { free(s->A); free(s->b); free(s->c); free(s->pos2var); free(s); } }

int td_get_entering(const td_state *s) {               /* LPState.java:274-285 */
  for (int i = 0; i < s->n; i++) if (dec_cmp(s->c[i], s->eps) > 0) return i;
  return -1;
}
int td_get_leaving(const td_state *s, int e) {          /* LPState.java:287-305 */
  int leaving = -1;
  dec min_slack = s->inf;
  for (int i = 0; i < s->m; i++) {
    dec aie = s->A[(size_t)i * s->n + e];
    dec slack = dec_cmp(aie, s->eps) < 0 ? s->inf : dec_div(s->b[i], aie);
    if (dec_cmp(slack, min_slack) < 0) { min_slack = slack; leaving = i; }
  }
  return leaving;
}
void td_pivot(td_state *s, int e, int l, int nthreads) { /* LPState.java:133-181 */
  const int n = s->n, m = s->m;
  dec *prow = s->A + (size_t)l * n;
  const dec p = prow[e];
  dec one = {1, 0, 0};
  prow[e] = dec_div(one, p);
  for (int j = 0; j < n; j++) if (j != e) prow[j] = dec_div(prow[j], p);
  s->b[l] = dec_div(s->b[l], p);
  const dec b_ent = s->b[l];
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(static) num_threads(nthreads) if (nthreads > 1)
  for (int i = 0; i < m; i++) {
    if (i == l) continue;
    dec *row = s->A + (size_t)i * n;
    const dec a = row[e];
    row[e] = dec_neg(dec_div(a, p));
    for (int j = 0; j < n; j++) if (j != e) row[j] = dec_sub(row[j], dec_mul(a, prow[j]));
    s->b[i] = dec_sub(s->b[i], dec_mul(a, b_ent));
  }
  const dec ce = s->c[e];
  s->v = dec_add(s->v, dec_mul(s->b[l], ce));
  s->c[e] = dec_neg(dec_div(ce, p));
  for (int j = 0; j < n; j++) if (j != e) s->c[j] = dec_sub(s->c[j], dec_mul(ce, prow[j]));
  int t = s->pos2var[e]; s->pos2var[e] = s->pos2var[n + l]; s->pos2var[n + l] = t;
}
int td_run(td_state *s, long max_pivots, int *log, long log_cap, long *npivots, int nthreads) {
  long k = 0; int status = TD_OPTIMAL;      /* LPSolver.java:101-112 */
  for (;;) {
    int e = td_get_entering(s);
    if (e == -1) break;
    int l = td_get_leaving(s, e);
    if (l == -1) { status = TD_UNBOUNDED; break; }
    if (max_pivots >= 0 && k >= max_pivots) { status = TD_PIVOT_CAP; break; }
    td_pivot(s, e, l, nthreads);
    if (log && k < log_cap) { log[2 * k] = e; log[2 * k + 1] = l; }
    k++;
  }
  if (npivots) *npivots = k;
  return status;
}
void td_read(const td_state *s, double *A, double *b, double *c, double *v, int *pos2var) {
  if (A) for (size_t k = 0; k < (size_t)s->m * s->n; k++) A[k] = dec_to_double(s->A[k]);
  if (b) for (int i = 0; i < s->m; i++) b[i] = dec_to_double(s->b[i]);
  if (c) for (int j = 0; j < s->n; j++) c[j] = dec_to_double(s->c[j]);
  if (v) *v = dec_to_double(s->v);
  if (pos2var) memcpy(pos2var, s->pos2var, sizeof(int) * (size_t)(s->m + s->n));
}
void td_v_string(const td_state *s, char *buf, int cap) { dec_to_string(s->v, buf, (size_t)cap); }
void td_cell_string(const td_state *s, int i, int j, char *buf, int cap) {
  dec d = (i < s->m) ? (j < s->n ? s->A[(size_t)i * s->n + j] : s->b[i]) : s->c[j];
  dec_to_string(d, buf, (size_t)cap);
}
