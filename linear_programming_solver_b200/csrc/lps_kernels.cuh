// Device kernels of the B200-native simplex pivot loop (sm_100a).
//
// Data layout in HBM: augmented tableau T = [A | b ; c | -v], (m+1) x ld doubles, row-major,
// ld = pitch (multiple of 16 doubles = 128 B), columns n+1..ld-1 are zero padding.
// One pivot = k_ratio -> k_scale_row -> k_update; everything a later kernel needs from an
// earlier one travels through `Ctl` and the small staging vectors (colbuf, bcol, rowbuf), so
// the host is never in the loop.
//
// Arithmetic contract (bit-identical to oracle/tier_f.c): every cell update is a separately
// rounded multiply and subtract (__dmul_rn/__dsub_rn, never an FMA) and every quotient is a
// true IEEE division (__ddiv_rn), mirroring LPState.java:139-177 where each BigDecimal
// multiply/subtract/divide is rounded on its own.
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

namespace lps {

constexpr int kNone = INT_MAX;  // "no index" sentinel for atomicMin targets

enum Verdict : int { kRunning = 0, kOptimal = 1, kUnbounded = 2, kPivotCap = 3, kZeroPivot = 4 };

// Loop-carried control block, one per handle, in device memory.
struct Ctl {
  int status;               // Verdict
  int e_cur;                // entering column of the pivot in flight / last executed
  int l_cur;                // leaving row of that pivot
  int e_next;               // entering column for the NEXT pivot (kNone = none); atomicMin target
  long long npivots;        // pivots executed since load
  long long pivot_limit;    // run stops (kPivotCap) when npivots == pivot_limit and a pivot is pending
  double p;                 // pivot element A[l][e] (old value)
  unsigned int ticket;      // last-block-done counter of k_ratio
  int q_leaving;            // result slot of the query-only ratio test (lps_get_leaving)
  int q_index;              // result slot of k_first_nonzero
  int pad_;
};

struct Cand {
  double slack;
  int row;
  int pad_;
};

__device__ __forceinline__ Cand cand_min(Cand a, Cand b) {
  // lexicographic (slack, row): the sequential scan of LPState.java:292-303 keeps the first
  // (lowest-index) row among equal minimal ratios because its comparison is strict.
  if (b.slack < a.slack || (b.slack == a.slack && b.row < a.row)) return b;
  return a;
}

__device__ __forceinline__ Cand warp_cand_min(Cand c) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    Cand o;
    o.slack = __shfl_xor_sync(0xffffffffu, c.slack, off);
    o.row = __shfl_xor_sync(0xffffffffu, c.row, off);
    o.pad_ = 0;
    c = cand_min(c, o);
  }
  return c;
}

__device__ __forceinline__ int warp_min_int(int v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}

// ---------------------------------------------------------------------------------------------
// run set-up: one thread
__global__ void k_begin_run(Ctl* ctl, long long max_pivots, int reset_next) {
  ctl->status = kRunning;
  ctl->pivot_limit = (max_pivots < 0) ? LLONG_MAX : ctl->npivots + max_pivots;
  ctl->ticket = 0;
  if (reset_next) ctl->e_next = kNone;
}

// LPState.getEntering (LPState.java:274-285): min{ j : c[j] > eps } via atomicMin on the index.
__global__ void k_first_positive(Ctl* ctl, const double* __restrict__ crow, int n, double eps) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int mine = (j < n && crow[j] > eps) ? j : kNone;
  mine = warp_min_int(mine);
  if ((threadIdx.x & 31) == 0 && mine != kNone) atomicMin(&ctl->e_next, mine);
}

// Gather column `e` (all m+1 rows, the objective row included) and column n (b) into the
// contiguous staging vectors.  e < 0 means "use ctl->e_next".
__global__ void k_extract(const Ctl* ctl, const double* __restrict__ T, long long ld, int m, int n,
                          int e_arg, double* col0, double* col1, double* bcol) {
  int e = (e_arg >= 0) ? e_arg : ctl->e_next;
  double* col = (ctl->npivots & 1) ? col1 : col0;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > m) return;
  if (e != kNone) col[i] = T[(long long)i * ld + e];
  bcol[i] = T[(long long)i * ld + n];
}

// ---------------------------------------------------------------------------------------------
// K1 — ratio test, LPState.getLeaving (LPState.java:287-305), plus the loop bookkeeping of
// LPSolver.simplex (LPSolver.java:101-108) in the last block to finish.
// mode 0: loop step (decides verdict, commits the pivot: log, positions, npivots)
// mode 1: query only (writes ctl->q_leaving)
__global__ void k_ratio(Ctl* ctl, const double* __restrict__ col0, const double* __restrict__ col1,
                        const double* __restrict__ bcol, int m, int n, double eps, double inf,
                        Cand* partials, int2* plog, long long log_cap, int* pos2var, int mode) {
  if (mode == 0 && ctl->status != kRunning) return;
  const long long np = ctl->npivots;
  const double* col = (np & 1) ? col1 : col0;
  Cand best;
  best.slack = inf;
  best.row = kNone;
  best.pad_ = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    double a = col[i];
    if (!(a < eps)) {                       // aie.compareTo(epsilon) < 0 -> INF   (:294-296)
      double s = __ddiv_rn(bcol[i], a);     // b[i].divide(aie, rounder)          (:297)
      if (s < best.slack) {                 // strict: first row wins ties         (:299)
        best.slack = s;
        best.row = i;
      }
    }
  }
  __shared__ Cand sh[32];
  __shared__ bool is_last;
  best = warp_cand_min(best);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  if (lane == 0) sh[warp] = best;
  __syncthreads();
  if (warp == 0) {
    Cand c;
    c.slack = inf; c.row = kNone; c.pad_ = 0;
    if (lane < nwarp) c = sh[lane];
    c = warp_cand_min(c);
    if (lane == 0) {
      partials[blockIdx.x] = c;
      __threadfence();
      unsigned int t = atomicAdd(&ctl->ticket, 1u);
      is_last = (t == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // last block: reduce the per-block partials (gridDim.x <= 1024)
  Cand c;
  c.slack = inf; c.row = kNone; c.pad_ = 0;
  for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) c = cand_min(c, partials[k]);
  c = warp_cand_min(c);
  if (lane == 0) sh[warp] = c;
  __syncthreads();
  if (warp == 0) {
    Cand d;
    d.slack = inf; d.row = kNone; d.pad_ = 0;
    if (lane < nwarp) d = sh[lane];
    d = warp_cand_min(d);
    if (lane == 0) {
      ctl->ticket = 0;
      int l = (d.row == kNone) ? -1 : d.row;
      if (mode == 1) {
        ctl->q_leaving = l;
      } else {
        int e = ctl->e_next;
        if (e == kNone) {                       // getEntering() == -1: optimal   LPSolver.java:101
          ctl->status = kOptimal;
          ctl->e_cur = -1;
          ctl->l_cur = -1;
        } else if (l < 0) {                     // getLeaving() == -1: unbounded  LPSolver.java:103
          ctl->status = kUnbounded;
          ctl->e_cur = e;
          ctl->l_cur = -1;
        } else if (np >= ctl->pivot_limit) {
          ctl->status = kPivotCap;
          ctl->e_cur = e;
          ctl->l_cur = l;
        } else {                                // commit pivot(e, l)
          ctl->e_cur = e;
          ctl->l_cur = l;
          ctl->p = col[l];
          plog[np % log_cap] = make_int2(e, l);
          int t = pos2var[e];                   // exchangeIndexes, LPState.java:311-320
          pos2var[e] = pos2var[n + l];
          pos2var[n + l] = t;
          ctl->npivots = np + 1;
          ctl->e_next = kNone;
        }
      }
    }
  }
}

// explicit LPState.pivot(e, l): commit without a ratio test (one thread)
__global__ void k_set_pivot(Ctl* ctl, const double* col0, const double* col1, int e, int l, int n,
                            int2* plog, long long log_cap, int* pos2var) {
  const long long np = ctl->npivots;
  const double* col = (np & 1) ? col1 : col0;
  double p = col[l];
  ctl->ticket = 0;
  ctl->e_cur = e;
  ctl->l_cur = l;
  if (p == 0.0) {            // BigDecimal.divide by zero throws ArithmeticException (LPState.java:139)
    ctl->status = kZeroPivot;
    return;
  }
  ctl->status = kRunning;
  ctl->pivot_limit = LLONG_MAX;
  ctl->p = p;
  plog[np % log_cap] = make_int2(e, l);
  int t = pos2var[e];
  pos2var[e] = pos2var[n + l];
  pos2var[n + l] = t;
  ctl->npivots = np + 1;
  ctl->e_next = kNone;
}

// ---------------------------------------------------------------------------------------------
// K2 — pivot row, LPState.java:137-146: r_e = 1/p, r_j = A[l][j]/p, b_l = b_l/p; written back
// into T[l] and staged in rowbuf.  Fused: the next entering column.  The new objective row is
// c'_j = c_j - c_e*r_j (c'_e = -(c_e/p)), an O(n) computation that does not need the O(mn) pass,
// so min{ j : c'_j > eps } (LPState.java:274-285 applied to the state AFTER this pivot) is known
// before k_update starts and k_update can emit that column as it streams by.
__global__ void k_scale_row(Ctl* ctl, double* __restrict__ T, long long ld, int m, int n,
                            double* __restrict__ rowbuf, const double* __restrict__ col0,
                            const double* __restrict__ col1, double eps) {
  if (ctl->status != kRunning) return;
  const int e = ctl->e_cur, l = ctl->l_cur;
  const double p = ctl->p;
  const double* col = ((ctl->npivots - 1) & 1) ? col1 : col0;
  const double ce = col[m];
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int mine = kNone;
  if (j < ld) {
    double r = 0.0;
    if (j <= n) {
      double* tl = T + (long long)l * ld;
      r = (j == e) ? __ddiv_rn(1.0, p) : __ddiv_rn(tl[j], p);
      tl[j] = r;
      if (j < n) {
        double cj = T[(long long)m * ld + j];
        double cn = (j == e) ? -__ddiv_rn(ce, p) : __dsub_rn(cj, __dmul_rn(ce, r));
        if (cn > eps) mine = j;
      }
    }
    rowbuf[j] = r;
  }
  mine = warp_min_int(mine);
  if ((threadIdx.x & 31) == 0 && mine != kNone) atomicMin(&ctl->e_next, mine);
}

// ---------------------------------------------------------------------------------------------
// K3 — the tableau update, LPState.java:150-178 on the augmented tableau (the objective row is
// row m: its formulas :170-178 are the same as an ordinary row's :157-164):
//   i != l:  T[i][e] <- -(a_i/p);  T[i][j] <- T[i][j] - a_i * r_j  (j != e)
// with a_i = old column e (colbuf) and r = new pivot row (rowbuf).  One 8-byte read and one
// 8-byte write per cell; 256-bit global accesses; the row slice r_j lives in registers.
// Fused: emits the post-update entering column of the NEXT pivot (ctl->e_next) and the
// post-update b column into contiguous vectors, so the next ratio test reads 16*m bytes
// instead of touching the tableau again.
struct __align__(32) D4 {
  double x, y, z, w;
};

__device__ __forceinline__ D4 ld256(const double* p) {
  D4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st256(double* p, const D4& v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z),
               "d"(v.w)
               : "memory");
}

template <int kThreads, int kRowsPerCta, int kUnroll, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
k_update(const Ctl* __restrict__ ctl, double* __restrict__ T, long long ld, int m, int n,
         const double* __restrict__ rowbuf, double* col0, double* col1,
         double* __restrict__ bcol) {
  if (ctl->status != kRunning) return;
  const int e = ctl->e_cur, l = ctl->l_cur, e2 = ctl->e_next;
  const double p = ctl->p;
  const long long np = ctl->npivots;
  const double* acol = ((np - 1) & 1) ? col1 : col0;  // this pivot's entering column (old values)
  double* ncol = (np & 1) ? col1 : col0;              // next pivot's entering column (new values)

  const long long j0 = ((long long)blockIdx.x * kThreads + threadIdx.x) * 4;
  if (j0 >= ld) return;
  const D4 r = *reinterpret_cast<const D4*>(rowbuf + j0);
  // which of my four lanes (if any) is the pivot column / next entering column / b column
  const int ke = (e >= j0 && e < j0 + 4) ? (int)(e - j0) : -1;
  const int k2 = (e2 != kNone && e2 >= j0 && e2 < j0 + 4) ? (int)(e2 - j0) : -1;
  const int kb = (n >= j0 && n < j0 + 4) ? (int)(n - j0) : -1;
  const bool special = (ke >= 0) | (k2 >= 0) | (kb >= 0);

  const int i_begin = blockIdx.y * kRowsPerCta;
  const int i_end = min(i_begin + kRowsPerCta, m + 1);
  double* base = T + j0;

  for (int i = i_begin; i < i_end; i += kUnroll) {
    D4 t[kUnroll];
    double a[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
      int ii = i + u;
      if (ii < i_end) {
        a[u] = acol[ii];
        t[u] = ld256(base + (long long)ii * ld);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
      int ii = i + u;
      if (ii < i_end) {
        D4 o;
        if (ii != l) {
          o.x = __dsub_rn(t[u].x, __dmul_rn(a[u], r.x));
          o.y = __dsub_rn(t[u].y, __dmul_rn(a[u], r.y));
          o.z = __dsub_rn(t[u].z, __dmul_rn(a[u], r.z));
          o.w = __dsub_rn(t[u].w, __dmul_rn(a[u], r.w));
          if (ke >= 0) {
            double q = -__ddiv_rn(a[u], p);
            if (ke == 0) o.x = q; else if (ke == 1) o.y = q; else if (ke == 2) o.z = q; else o.w = q;
          }
          st256(base + (long long)ii * ld, o);
        } else {
          o = r;  // the pivot row was already rewritten by k_scale_row
        }
        if (special) {
          if (k2 >= 0) ncol[ii] = (k2 == 0) ? o.x : (k2 == 1) ? o.y : (k2 == 2) ? o.z : o.w;
          if (kb >= 0) bcol[ii] = (kb == 0) ? o.x : (kb == 1) ? o.y : (kb == 2) ? o.z : o.w;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// loaders / utilities
__global__ void k_set_column(double* T, long long ld, int m, int col, const double* __restrict__ src,
                             double constant, int use_constant) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) T[(long long)i * ld + col] = use_constant ? constant : src[i];
}

__global__ void k_gather_column(const double* __restrict__ T, long long ld, int m, int col,
                                double* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) dst[i] = T[(long long)i * ld + col];
}

__global__ void k_iota(int* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

// performDegeneratePivot's scan (LPSolver.java:185-191): first j with |A[row][j]| > eps
__global__ void k_first_nonzero(Ctl* ctl, const double* __restrict__ row, int n, double eps) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int mine = (j < n && fabs(row[j]) > eps) ? j : kNone;
  mine = warp_min_int(mine);
  if ((threadIdx.x & 31) == 0 && mine != kNone) atomicMin(&ctl->q_index, mine);
}

// restoreInitialLP's column removal (LPSolver.java:205-211) in place: every row shifts its
// cells j+1..n (b included) one place left.  One CTA per row; chunked so that a chunk's reads
// complete before its writes.
__global__ void k_drop_column(double* T, long long ld, int m, int n, int jdrop) {
  extern __shared__ double chunk[];
  for (int i = blockIdx.x; i <= m; i += gridDim.x) {
    double* row = T + (long long)i * ld;
    for (int base = jdrop; base < n; base += blockDim.x) {
      int j = base + threadIdx.x;             // destination column
      double v = 0.0;
      if (j < n) v = row[j + 1];              // source j+1 <= n
      __syncthreads();
      if (j < n) row[j] = v;
      __syncthreads();
    }
    if (threadIdx.x == 0) row[n] = 0.0;       // keep the padding zero
    __syncthreads();
  }
}

struct ObjOp {
  int kind;
  int index;
  double coef;
};

// restoreInitialLP's objective rebuild (LPSolver.java:213-233): one thread per column applies
// the ops in the given order with separately rounded multiply and add; thread n does v.
__global__ void k_rebuild_objective(double* T, long long ld, int m, int n,
                                    const ObjOp* __restrict__ ops, int nops) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > n) return;
  double acc = 0.0;
  if (j < n) {
    for (int k = 0; k < nops; k++) {
      ObjOp op = ops[k];
      if (op.kind == 0) {
        double a = T[(long long)op.index * ld + j];
        acc = __dadd_rn(acc, __dmul_rn(-a, op.coef));       // :226-227
      } else if (op.index == j) {
        acc = __dadd_rn(acc, op.coef);                      // :231
      }
    }
    T[(long long)m * ld + j] = acc;
  } else {
    for (int k = 0; k < nops; k++) {
      ObjOp op = ops[k];
      if (op.kind == 0) acc = __dadd_rn(acc, __dmul_rn(T[(long long)op.index * ld + n], op.coef));  // :223
    }
    T[(long long)m * ld + n] = -acc;                         // corner holds -v
  }
}

// ---------------------------------------------------------------------------------------------
// counter-based generator of the synthetic LPs (kernels in lps_sharded.cuh)
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
__device__ __forceinline__ double synth_u(unsigned long long seed, unsigned long long k) {
  unsigned long long h = splitmix64(seed ^ (k * 0x9E3779B97F4A7C15ULL));
  return (double)((h >> 44) + 1ULL) * (1.0 / 1048576.0);
}

}  // namespace lps
