// Blocked pivot loop: S pivots per pass over the tableau instead of one.
//
// The rank-1 update of LPState.pivot (LPState.java:150-178) moves 16 bytes per cell for two
// flops, so a pivot-per-pass loop is pinned to HBM bandwidth / 16(m+1)(n+1).  But the only values
// pivot t needs from the tableau BEFORE deciding the next pivot are one column (the entering one,
// for the ratio test of LPState.getLeaving, :287-305), the b column, one row (the leaving one, to
// scale it, :137-146) and the objective row (for LPState.getEntering, :274-285).  So the loop
// keeps the tableau T as it was at the start of a block, plus the pending pivots' columns a_u
// (old entering column), rows r_u (new, scaled pivot row), indices (e_u, l_u) and pivot elements
// p_u, and evaluates any cell it needs lazily by replaying the pending pivots on it:
//
//     x <- T[i][j];  for u = 0..t-1:   i == l_u  ->  x = r_u[j]
//                                      j == e_u  ->  x = -(a_u[i] / p_u)
//                                      else      ->  x = x - a_u[i] * r_u[j]
//
// which is literally the sequence of separately rounded operations the pivot-per-pass loop (and
// the reference) applies to that cell, so every value — hence every comparison, hence the pivot
// sequence — is bit-identical.  After S pivots (or at a verdict) kb_flush replays the pending
// pivots on every cell in ONE pass: one HBM read and one HBM write of the tableau per S pivots.
//
//   kb_col    entering column + b column of the current state (two strided gathers + replay),
//             ratio test, per-rank candidate pushed to every peer            [O(m t) work]
//   kb_row    winner of the candidates; owner replays + scales the leaving row and stores it
//             into every rank's row store (chunk flags); everyone replays the objective row and
//             finds the next entering column; last CTA commits the pivot     [O(n t) work]
//   kb_flush  T <- T with all pending pivots applied                         [O(m n t) flops, 16 B/cell]
//
// Row sharding is the same as lps_sharded.cuh (rows local, objective row replicated, candidates
// and scaled rows by stores into peer memory); a single GPU is world == 1 talking to itself.
#pragma once
#include "lps_sharded.cuh"

namespace lps {

// one pending pivot replayed on one cell
__device__ __forceinline__ double replay(double x, bool pivot_row, bool pivot_col, double a, double r,
                                         double p) {
  if (pivot_row) return r;                       // LPState.java:137-146 (row l is the scaled row)
  if (pivot_col) return -__ddiv_rn(a, p);        // :157 / :172
  return __dsub_rn(x, __dmul_rn(a, r));          // :162-164 / :177
}

constexpr int kColThreads = 128;

// K1 (blocked): column e and column n of the CURRENT state for the local rows (objective row
// included), ratio test, candidate push.  Acols[t] <- column e.
__global__ void __launch_bounds__(kColThreads)
kb_col(CtlS* ctl, const double* __restrict__ T, long long ld, int mloc, int n, int row0,
       double* __restrict__ Acols, long long apitch, const double* Rrows, double eps, double inf,
       Cand* partials, Peers peers, int rank, int world) {
  if (ctl->base.status != kRunning) return;
  const long long np = ctl->base.npivots;
  const unsigned int seq = (unsigned int)(np + 1);
  const int t = ctl->blk_pending;
  const int e = ctl->e_nx[seq & 1];
  __shared__ double s_re[kMaxBlock], s_rn[kMaxBlock], s_p[kMaxBlock];
  __shared__ int s_l[kMaxBlock], s_e[kMaxBlock];
  if ((int)threadIdx.x < t) {
    const int u = threadIdx.x;
    s_l[u] = ctl->blk_l[u];
    s_e[u] = ctl->blk_e[u];
    s_p[u] = ctl->blk_p[u];
    s_rn[u] = Rrows[(long long)u * ld + n];
    s_re[u] = (e != kNone) ? Rrows[(long long)u * ld + e] : 0.0;
  }
  __syncthreads();
  Cand best;
  best.slack = inf; best.row = kNone; best.pad_ = 0;
  if (e != kNone) {
    double* acol = Acols + (long long)t * apitch;
    for (int i = blockIdx.x * kColThreads + threadIdx.x; i <= mloc; i += gridDim.x * kColThreads) {
      double xe = T[(long long)i * ld + e];
      double xb = T[(long long)i * ld + n];
      for (int u = 0; u < t; u++) {
        const double a = Acols[(long long)u * apitch + i];
        const bool prow = (i == s_l[u]);
        xe = replay(xe, prow, e == s_e[u], a, s_re[u], s_p[u]);
        xb = replay(xb, prow, false, a, s_rn[u], s_p[u]);
      }
      acol[i] = xe;
      if (i < mloc && !(xe < eps)) {              // aie.compareTo(epsilon) < 0 -> INF   (:294-296)
        double s = __ddiv_rn(xb, xe);              // b[i].divide(aie, rounder)          (:297)
        if (s < best.slack) { best.slack = s; best.row = i; }   // strict: first row wins ties (:299)
      }
    }
  }
  __shared__ Cand sh[kColThreads / 32];
  __shared__ bool is_last;
  best = warp_cand_min(best);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    Cand c = sh[0];
#pragma unroll
    for (int w = 1; w < kColThreads / 32; w++) c = cand_min(c, sh[w]);
    partials[blockIdx.x] = c;
    __threadfence();
    unsigned int tk = atomicAdd(&ctl->base.ticket, 1u);
    is_last = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last || warp != 0) return;
  __threadfence();
  Cand d;
  d.slack = inf; d.row = kNone; d.pad_ = 0;
  for (int k = lane; k < (int)gridDim.x; k += 32) {
    Cand o;
    o.slack = __ldcg(&partials[k].slack);
    o.row = __ldcg(&partials[k].row);
    o.pad_ = 0;
    d = cand_min(d, o);
  }
  d = warp_cand_min(d);
  if (lane == 0) {
    ctl->base.ticket = 0;
    ctl->e_nx[(seq & 1) ^ 1] = kNone;   // atomicMin target of this pivot's kb_row
  }
  if (lane < world) {                   // one lane per peer (own mailbox included)
    PeerCand* dst = &peers.blk[lane]->cand[seq & 1][rank];
    dst->slack = d.slack;
    dst->p = (d.row == kNone) ? 0.0 : __ldcg(Acols + (long long)t * apitch + d.row);
    dst->row = (d.row == kNone) ? kNone : row0 + d.row;
    __threadfence_system();
    st_release_sys(&dst->seq, seq);
  }
}

// K2 (blocked): candidates -> winner; the owner replays and scales its chunk of the leaving row
// and stores it into every rank's row store; every rank replays its chunk of the objective row,
// applies this pivot to it on the fly and looks for the next entering column; the last CTA
// commits the pivot into the pending block.
__global__ void __launch_bounds__(kChunk)
kb_row(CtlS* ctl, const double* __restrict__ T, long long ld, int mloc, int n, int row0, int row1,
       const double* __restrict__ Acols, long long apitch, double eps, double inf, Peers peers,
       int rank, int world, int2* plog, long long log_cap, int* pos2var) {
  if (ctl->base.status != kRunning) return;
  const long long np = ctl->base.npivots;
  const unsigned int seq = (unsigned int)(np + 1);
  const int par = seq & 1;
  const int t = ctl->blk_pending;
  CommBlock* mine = peers.blk[rank];
  __shared__ int s_lw, s_verdict;
  __shared__ double s_pw;
  __shared__ double s_al[kMaxBlock], s_am[kMaxBlock], s_p[kMaxBlock];
  __shared__ int s_l[kMaxBlock], s_e[kMaxBlock];
  const int e = ctl->e_nx[par];
  if (threadIdx.x < 32) {
    bool ok = true;
    PeerCand c;
    c.slack = inf; c.row = kNone; c.p = 0.0;
    if ((int)threadIdx.x < world) {
      const PeerCand* src = &mine->cand[par][threadIdx.x];
      ok = spin_until(&src->seq, seq);
      c.slack = ld_volatile_f64(&src->slack);
      c.p = ld_volatile_f64(&src->p);
      c.row = ld_volatile_s32(&src->row);
    }
    ok = __all_sync(0xffffffffu, ok);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      double os = __shfl_xor_sync(0xffffffffu, c.slack, off);
      double op = __shfl_xor_sync(0xffffffffu, c.p, off);
      int orow = __shfl_xor_sync(0xffffffffu, c.row, off);
      if (os < c.slack || (os == c.slack && orow < c.row)) { c.slack = os; c.p = op; c.row = orow; }
    }
    if (threadIdx.x == 0) {
      int verdict = kRunning;
      if (!ok) verdict = kCommTimeout;
      else if (e == kNone) verdict = kOptimal;             // getEntering() == -1   LPSolver.java:101
      else if (c.row == kNone) verdict = kUnbounded;       // getLeaving() == -1    LPSolver.java:103
      else if (np >= ctl->base.pivot_limit) verdict = kPivotCap;
      s_verdict = verdict;
      s_lw = (c.row == kNone) ? -1 : c.row;
      s_pw = c.p;
    }
  }
  __syncthreads();
  const int verdict = s_verdict;
  const int l = s_lw;
  const double p = s_pw;
  const bool i_own = (l >= row0 && l < row1);
  const int lloc = l - row0;
  if (verdict == kRunning && (int)threadIdx.x < t) {
    const int u = threadIdx.x;
    s_l[u] = ctl->blk_l[u];
    s_e[u] = ctl->blk_e[u];
    s_p[u] = ctl->blk_p[u];
    s_al[u] = i_own ? Acols[(long long)u * apitch + lloc] : 0.0;
    s_am[u] = Acols[(long long)u * apitch + mloc];
  }
  __syncthreads();
  const int j = blockIdx.x * kChunk + threadIdx.x;
  int mine_next = kNone;
  if (verdict == kRunning) {
    const double* rows = peers.rowbuf[rank];      // this rank's copy of the pending rows
    double r = 0.0;
    if (i_own) {
      if (j < ld) {
        if (j <= n) {
          double x = T[(long long)lloc * ld + j];
          for (int u = 0; u < t; u++)
            x = replay(x, lloc == s_l[u], j == s_e[u], s_al[u], rows[(long long)u * ld + j], s_p[u]);
          r = (j == e) ? __ddiv_rn(1.0, p) : __ddiv_rn(x, p);        // LPState.java:139-146
        }
        for (int k = 0; k < world; k++) (peers.rowbuf[k] + (long long)t * ld)[j] = r;
      }
      __threadfence_system();
      __syncthreads();
      if ((int)threadIdx.x < world && (int)threadIdx.x != rank)
        st_release_sys(&peers.blk[threadIdx.x]->row_flag[par][blockIdx.x], seq);
    } else {
      __shared__ bool s_ok;
      if (threadIdx.x == 0) s_ok = spin_until(&mine->row_flag[par][blockIdx.x], seq);
      __syncthreads();
      if (!s_ok) {
        if (threadIdx.x == 0) ctl->base.status = kCommTimeout;
        return;
      }
      if (j < ld) r = ld_volatile_f64(rows + (long long)t * ld + j);
    }
    if (j < n) {
      double xc = T[(long long)mloc * ld + j];
      for (int u = 0; u < t; u++)
        xc = replay(xc, false, j == s_e[u], s_am[u], rows[(long long)u * ld + j], s_p[u]);
      const double ce = Acols[(long long)t * apitch + mloc];
      double cn = (j == e) ? -__ddiv_rn(ce, p) : __dsub_rn(xc, __dmul_rn(ce, r));   // :170-178
      if (cn > eps) mine_next = j;
    }
    mine_next = warp_min_int(mine_next);
    if ((threadIdx.x & 31) == 0 && mine_next != kNone) atomicMin(&ctl->e_nx[par ^ 1], mine_next);
  }
  // last CTA to finish commits (every CTA has read npivots / pending / e_nx[par] by then)
  __shared__ bool is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int tk = atomicAdd(&ctl->ticket2, 1u);
    is_last = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last || threadIdx.x != 0) return;
  ctl->ticket2 = 0;
  if (verdict != kRunning) {
    ctl->base.status = verdict;
    ctl->base.e_cur = (e == kNone) ? -1 : e;
    ctl->base.l_cur = (verdict == kPivotCap) ? l : -1;
    return;
  }
  ctl->base.e_cur = e;
  ctl->base.l_cur = l;
  ctl->base.p = p;
  ctl->owner = i_own ? rank : -1;
  ctl->blk_e[t] = e;
  ctl->blk_l[t] = i_own ? lloc : -1;
  ctl->blk_p[t] = p;
  ctl->blk_pending = t + 1;
  plog[np % log_cap] = make_int2(e, l);
  int tmp = pos2var[e];                                   // exchangeIndexes, LPState.java:311-320
  pos2var[e] = pos2var[n + l];
  pos2var[n + l] = tmp;
  ctl->base.npivots = np + 1;
}

// K3 (blocked): apply all pending pivots to every local cell (objective row included) in one
// pass: one 256-bit load and one 256-bit store per four cells, 2t flops per cell in between.
//
// One persistent CTA per SM, kLanes x 128 threads.  Work is cut into chunks of kCH rows x 512
// columns, numbered row-band-major (consecutive chunks are neighbouring strips of the same rows, so
// the CTAs of the grid sweep the tableau as one band: DRAM pages stay open) and claimed from an
// atomic counter two chunks ahead.  Each 128-thread lane owns whole groups of kU rows (kU 256-bit
// loads in flight per thread).  The operands of the replay are double-buffered in shared memory
// and arrive by cp.async while the previous chunk is being computed:
//   s_r [2][t][512]   the strip's slice of the pending rows
//   s_a [2][t][kCH]   the chunk's slice of the pending columns
// so the inner loop is LDS + DMUL/DADD only.  A group of rows that holds no pending pivot row, in a
// thread that holds no pending pivot column, runs branch-free; anything else takes the generic
// replay() path row by row.
constexpr int kFlushThreads = 128;
constexpr int kStripCols = 4 * kFlushThreads;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned int sa = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int kLanes, int kU, int kG, bool kPre>
__global__ void __launch_bounds__(kFlushThreads * kLanes, 1)
kb_flush(CtlS* ctl, double* __restrict__ T, long long ld, int mloc, const double* __restrict__ Acols,
         long long apitch, const double* __restrict__ Rrows) {
  constexpr int kThreads = kFlushThreads * kLanes;
  constexpr int kCH = kU * kLanes * kG;            // rows per chunk
  static_assert(kU % 2 == 0, "16-byte copies of the pending columns");
  const int t = ctl->blk_pending;
  if (t == 0) return;
  extern __shared__ __align__(32) double smem[];
  double* const s_r = smem;                                      // [2][t][kStripCols]
  double* const s_a = smem + (size_t)2 * t * kStripCols;         // [2][t][kCH]
  __shared__ double s_p[kMaxBlock];
  __shared__ int s_l[kMaxBlock], s_e[kMaxBlock];
  __shared__ long long s_claim[2], s_first[3];
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid / kFlushThreads, ltid = tid % kFlushThreads;
  if (tid < t) {
    s_l[tid] = ctl->blk_l[tid];
    s_e[tid] = ctl->blk_e[tid];
    s_p[tid] = ctl->blk_p[tid];
  }
  const int nstrips = (int)((ld + kStripCols - 1) / kStripCols);
  const int nrb = (mloc + 1 + kCH - 1) / kCH;                    // chunks per strip
  const long long nchunks = (long long)nstrips * nrb;
  unsigned long long* const queue = &ctl->blk_queue;
  if (tid == 0) {
    s_first[0] = (long long)atomicAdd(queue, 1ull);
    s_first[1] = (long long)atomicAdd(queue, 1ull);
    s_first[2] = (long long)atomicAdd(queue, 1ull);
  }
  __syncthreads();
  long long cur = s_first[0], nxt = s_first[1], nn = s_first[2];

  // chunk c's operand slices -> buffer `buf`
  auto prefetch = [&](long long c, int buf) {
    const long long jb = (c % nstrips) * kStripCols;
    const int i0 = (int)(c / nstrips) * kCH;
    double* dr = s_r + (size_t)buf * t * kStripCols;
    for (int idx = tid; idx < t * (kStripCols / 2); idx += kThreads) {
      const int u = idx / (kStripCols / 2), q = idx % (kStripCols / 2);
      // two planes of 16-byte pairs (x,y | z,w per thread) so that the 128-bit LDS are conflict-free
      if (jb + 2 * q < ld)
        cp_async16(dr + u * kStripCols + (q & 1) * (kStripCols / 2) + (q >> 1) * 2,
                   Rrows + (long long)u * ld + jb + 2 * q);
    }
    double* da = s_a + (size_t)buf * t * kCH;
    for (int idx = tid; idx < t * (kCH / 2); idx += kThreads) {
      const int u = idx / (kCH / 2), q = idx % (kCH / 2);
      cp_async16(da + u * kCH + 2 * q, Acols + (long long)u * apitch + i0 + 2 * q);
    }
    cp_async_commit();
  };

  if (cur < nchunks) prefetch(cur, 0);
  int buf = 0;
  bool first = true;
  while (cur < nchunks) {
    cp_async_wait_all();
    __syncthreads();   // this chunk's operands have landed; everyone is done with the other buffer
    if (!first) nn = s_claim[buf ^ 1];
    first = false;
    if (nxt < nchunks) prefetch(nxt, buf ^ 1);
    if (tid == 0) s_claim[buf] = (long long)atomicAdd(queue, 1ull);   // read after the next barrier

    const long long j0 = (cur % nstrips) * kStripCols + 4 * ltid;
    if (j0 < ld) {
      const int i0 = (int)(cur / nstrips) * kCH;
      const int i_end = min(i0 + kCH, mloc + 1);
      const double* sr = s_r + (size_t)buf * t * kStripCols + 2 * ltid;
      const double* sa = s_a + (size_t)buf * t * kCH;
      unsigned int cmask = 0;   // pending pivots whose entering column is one of my four
      for (int u = 0; u < t; u++)
        if (s_e[u] >= j0 && s_e[u] < j0 + 4) cmask |= 1u << u;
      double* base = T + j0;
      auto rvec = [&](int u) {
        const double2 lo = *reinterpret_cast<const double2*>(sr + u * kStripCols);
        const double2 hi = *reinterpret_cast<const double2*>(sr + u * kStripCols + kStripCols / 2);
        D4 r;
        r.x = lo.x; r.y = lo.y; r.z = hi.x; r.w = hi.y;
        return r;
      };
      auto load_group = [&](D4 (&x)[kU], int i) {
#pragma unroll
        for (int k = 0; k < kU; k++)
          if (i + k < i_end) x[k] = ld256(base + (long long)(i + k) * ld);
      };
      // replay the pending pivots on one group of rows held in registers, then store it
      auto finish_group = [&](D4 (&x)[kU], int i, int g) {
        // a full group without a pending pivot row, in a thread without a pending pivot column?
        bool plain = (cmask == 0) && (i + kU <= i_end);
        for (int u = 0; u < t; u++) plain &= !(s_l[u] >= i && s_l[u] < i + kU);
        if (plain) {
#pragma unroll 2
          for (int u = 0; u < t; u++) {
            const D4 r = rvec(u);
            double av[kU];
#pragma unroll
            for (int q = 0; q < kU / 2; q++) {
              const double2 v = *reinterpret_cast<const double2*>(sa + u * kCH + g * kU + 2 * q);
              av[2 * q] = v.x;
              av[2 * q + 1] = v.y;
            }
#pragma unroll
            for (int k = 0; k < kU; k++) {
              x[k].x = __dsub_rn(x[k].x, __dmul_rn(av[k], r.x));
              x[k].y = __dsub_rn(x[k].y, __dmul_rn(av[k], r.y));
              x[k].z = __dsub_rn(x[k].z, __dmul_rn(av[k], r.z));
              x[k].w = __dsub_rn(x[k].w, __dmul_rn(av[k], r.w));
            }
          }
#pragma unroll
          for (int k = 0; k < kU; k++) st256(base + (long long)(i + k) * ld, x[k]);
        } else {
          for (int u = 0; u < t; u++) {
            const D4 r = rvec(u);
            const int lu = s_l[u];
            const int ce = ((cmask >> u) & 1u) ? (int)(s_e[u] - j0) : -1;
            const double pu = s_p[u];
#pragma unroll
            for (int k = 0; k < kU; k++) {
              if (i + k < i_end) {
                const double a = sa[u * kCH + g * kU + k];
                const bool prow = (i + k == lu);
                x[k].x = replay(x[k].x, prow, ce == 0, a, r.x, pu);
                x[k].y = replay(x[k].y, prow, ce == 1, a, r.y, pu);
                x[k].z = replay(x[k].z, prow, ce == 2, a, r.z, pu);
                x[k].w = replay(x[k].w, prow, ce == 3, a, r.w, pu);
              }
            }
          }
#pragma unroll
          for (int k = 0; k < kU; k++)
            if (i + k < i_end) st256(base + (long long)(i + k) * ld, x[k]);
        }
      };
      if (kPre) {
        // the next group's rows are in flight while this group is computed
        D4 x[kU], xn[kU];
        int g = lane, i = i0 + g * kU;
        if (i < i_end) load_group(x, i);
        while (i < i_end) {
          const int gn = g + kLanes, in = i0 + gn * kU;
          const bool more = (gn < kCH / kU) && (in < i_end);
          if (more) load_group(xn, in);
          finish_group(x, i, g);
          if (!more) break;
#pragma unroll
          for (int k = 0; k < kU; k++) x[k] = xn[k];
          g = gn;
          i = in;
        }
      } else {
        for (int g = lane; g < kCH / kU; g += kLanes) {
          const int i = i0 + g * kU;
          if (i >= i_end) break;
          D4 x[kU];
          load_group(x, i);
          finish_group(x, i, g);
        }
      }
    }
    cur = nxt;
    nxt = nn;
    buf ^= 1;
  }
  // last CTA of the grid retires the block
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    unsigned int tk = atomicAdd(&ctl->blk_ticket, 1u);
    s_last = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && tid == 0) {
    ctl->blk_ticket = 0;
    ctl->blk_queue = 0;
    ctl->blk_pending = 0;
  }
}

}  // namespace lps
