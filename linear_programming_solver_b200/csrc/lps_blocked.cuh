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
#include "lps_loop.cuh"

namespace lps {

// one pending pivot replayed on one cell
__device__ __forceinline__ double replay(double x, bool pivot_row, bool pivot_col, double a, double r,
                                         double p) {
  if (pivot_row) return r;                       // LPState.java:137-146 (row l is the scaled row)
  if (pivot_col) return -__ddiv_rn(a, p);        // :157 / :172
  return __dsub_rn(x, __dmul_rn(a, r));          // :162-164 / :177
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned int sa = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// L2 eviction priority (createpolicy) for the tableau stream of the out-of-place pass: it goes through the 126 MB
// L2 once per pass while the panel re-reads a few MB of pending columns / rows per pivot.  Off by default (env
// LPS_L2_HINTS): neither this nor evict_last on the panel's operands nor an L2 persisting window measured a gain
// (profiles/r02_summary.md).
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ D4 ld256_hint(const double* p, unsigned long long pol) {
  D4 v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w)
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st256_hint(double* p, const D4& v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z),
               "d"(v.w), "l"(pol)
               : "memory");
}
constexpr int kColThreads = 128;
constexpr int kPanelMax = 20;
// The panel kernels replay up to kPanelMax pending pivots on a handful of cells per thread.  A
// pending pivot whose leaving row / entering column IS the cell's row / column overwrites the
// cell, so everything before the last such pivot is irrelevant: find it in a short rolled loop
// (the only place a division can occur), then run the plain mul/sub chain from there, unrolled
// and predicated.  Same operations on the surviving chain, so same bits — and a fraction of the
// code size of an unrolled chain of three-way replay() calls (these kernels run once per pivot,
// mostly out of a cold instruction cache).

// cell (i, e) and cell (i, n) of one row: av[u] = a_u[i] (registers), re/rn = r_u[e], r_u[n]
__device__ __forceinline__ void replay_row_pair(double& xe, double& xb, const double (&av)[kPanelMax],
                                                const double* __restrict__ acols_i, long long apitch, int t,
                                                int i, int e, const int* s_l, const int* s_e, const double* s_p,
                                                const double* s_re, const double* s_rn) {
  int ue = 0, ub = 0;
  for (int u = 0; u < t; u++) {
    if (i == s_l[u]) {                     // row i is pending pivot u's (scaled) leaving row
      xe = s_re[u];
      xb = s_rn[u];
      ue = ub = u + 1;
    } else if (e == s_e[u]) {              // column e was pending pivot u's entering column
      xe = -__ddiv_rn(acols_i[(long long)u * apitch], s_p[u]);
      ue = u + 1;
    }
  }
#pragma unroll
  for (int u = 0; u < kPanelMax; u++) {
    if (u < t) {
      if (u >= ue) xe = __dsub_rn(xe, __dmul_rn(av[u], s_re[u]));
      if (u >= ub) xb = __dsub_rn(xb, __dmul_rn(av[u], s_rn[u]));
    }
  }
}

// cell (row, j) of one column: rv[u] = r_u[j] (registers), s_a[u] = a_u[row]; row_is[u] != 0 when
// the row is pending pivot u's leaving row
__device__ __forceinline__ double replay_col_cell(double x, const double (&rv)[kPanelMax],
                                                  const double* rows_j, long long ld, int t, int j,
                                                  int row_local, const int* s_l, const int* s_e,
                                                  const double* s_p, const double* s_a) {
  int us = 0;
  for (int u = 0; u < t; u++) {
    if (row_local >= 0 && row_local == s_l[u]) {   // the row was pending pivot u's leaving row
      x = rows_j[(long long)u * ld];
      us = u + 1;
    } else if (j == s_e[u]) {                      // column j was pending pivot u's entering column
      x = -__ddiv_rn(s_a[u], s_p[u]);
      us = u + 1;
    }
  }
#pragma unroll
  for (int u = 0; u < kPanelMax; u++)
    if (u < t && u >= us) x = __dsub_rn(x, __dmul_rn(s_a[u], rv[u]));
  return x;
}
   // most pending pivots the kernels are built for (lps_create clamps block_pivots)

// K1 (blocked): column e and column n of the CURRENT state for the local rows (objective row
// included), ratio test, candidate push.  Acols[t] <- column e.
__global__ void __launch_bounds__(kColThreads)
kb_col(CtlS* ctl, const double* __restrict__ T, long long ld, int mloc, int n, int row0,
       double* __restrict__ Acols, long long apitch, const double* Rrows, double eps, double inf,
       Cand* partials, Peers peers, int rank, int world) {
  if (ctl->base.status != kRunning) return;
  const long long np = ctl->base.npivots;
  const unsigned int seq = (unsigned int)(np + 1);
  const int t = ctl->blk_pend[0];
  const int e = ctl->e_nx[seq & 1];
  __shared__ double s_re[kMaxBlock], s_rn[kMaxBlock], s_p[kMaxBlock];
  __shared__ int s_l[kMaxBlock], s_e[kMaxBlock];
  if ((int)threadIdx.x < t) {
    const int u = threadIdx.x;
    s_l[u] = ctl->blk_l2[0][u];
    s_e[u] = ctl->blk_e2[0][u];
    s_p[u] = ctl->blk_p2[0][u];
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
      // all operand loads first (one memory latency, not one per pending pivot), then the replay
      double av[kPanelMax];
#pragma unroll
      for (int u = 0; u < kPanelMax; u++)
        if (u < t) av[u] = Acols[(long long)u * apitch + i];
      replay_row_pair(xe, xb, av, Acols + i, apitch, t, i, e, s_l, s_e, s_p, s_re, s_rn);
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
  const int t = ctl->blk_pend[0];
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
    s_l[u] = ctl->blk_l2[0][u];
    s_e[u] = ctl->blk_e2[0][u];
    s_p[u] = ctl->blk_p2[0][u];
    s_al[u] = i_own ? Acols[(long long)u * apitch + lloc] : 0.0;
    s_am[u] = Acols[(long long)u * apitch + mloc];
  }
  __syncthreads();
  const int j = blockIdx.x * kChunk + threadIdx.x;
  int mine_next = kNone;
  if (verdict == kRunning) {
    const double* rows = peers.rowbuf[rank];      // this rank's copy of the pending rows
    double r = 0.0;
    // all operand loads first (one memory latency, not one per pending pivot), then the replays
    double rv[kPanelMax];
    double xc = 0.0;
    if (j < ld) {
#pragma unroll
      for (int u = 0; u < kPanelMax; u++)
        if (u < t) rv[u] = rows[(long long)u * ld + j];
      if (j < n) xc = T[(long long)mloc * ld + j];
    }
    if (i_own) {
      if (j < ld) {
        if (j <= n) {
          double x = T[(long long)lloc * ld + j];
          x = replay_col_cell(x, rv, rows + j, ld, t, j, lloc, s_l, s_e, s_p, s_al);
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
      xc = replay_col_cell(xc, rv, rows + j, ld, t, j, -1, s_l, s_e, s_p, s_am);
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
  ctl->blk_e2[0][t] = e;
  ctl->blk_l2[0][t] = i_own ? lloc : -1;
  ctl->blk_p2[0][t] = p;
  ctl->blk_pend[0] = t + 1;
  plog[np % log_cap] = make_int2(e, l);
  int tmp = pos2var[e];                                   // exchangeIndexes, LPState.java:311-320
  pos2var[e] = pos2var[n + l];
  pos2var[n + l] = tmp;
  ctl->base.npivots = np + 1;
}

// Persistent panel: ONE cooperative launch runs the column step and the row step of every pivot
// of a block (until the block is full, a verdict, or the pivot cap).  One CTA per SM; rows (phase
// A) and columns (phase B) are dealt to CTAs statically, so a thread re-reads only pending-column
// / pending-row entries its own CTA wrote.  The two global decisions of a pivot — the ratio-test
// winner and the next entering column — are all-gathers: every CTA writes its partial result into
// its own slot, the grid meets in panel_sync (ticket + one go word), then thread k of every CTA
// reads CTA k's slot and the CTA reduces the 148 values itself.  Values and decisions are the same
// as kb_col / kb_row, bit for bit.
constexpr int kPanelThreads = 384;
constexpr int kSlotStride = 128 / sizeof(PeerCand);             // slots are polled by every CTA: one L2 line each
constexpr int kMinStride = 128 / sizeof(unsigned long long);

// Flag-in-data packets for the cross-rank exchange of kb_panel: a double travels as one 16-byte store
// {lo32, tag, hi32, tag}.  Each 8-byte half carries the tag, so only 8-byte store atomicity is assumed;
// the receiver polls the packet itself until both tags match.  No flag word, no system-scope fence:
// an exchange costs one NVLink traversal.  Tags are pivot numbers (never 0, the cleared state); slots
// are double-buffered by pivot parity, so a stale packet carries tag - 2.
struct __align__(16) LLPacket {
  unsigned int lo, tag0, hi, tag1;
};
__device__ __forceinline__ void ll_store(LLPacket* dst, double v, unsigned int tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"((unsigned int)b), "r"(tag),
               "r"((unsigned int)(b >> 32)), "r"(tag)
               : "memory");
}
// false on timeout
__device__ __forceinline__ bool ll_load(const LLPacket* src, unsigned int tag, double& v) {
  unsigned int lo, t0, hi, t1;
  unsigned int spins = 0;
  unsigned long long start = 0;
  for (;;) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(t0), "=r"(hi), "=r"(t1) : "l"(src) : "memory");
    if (t0 == tag && t1 == tag) break;
    if ((++spins & 1023u) == 0) {
      const unsigned long long now = globaltimer_ns();
      if (start == 0) start = now;
      else if (now - start > kSpinTimeoutNs) return false;
    }
  }
  v = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
  return true;
}

struct PanelArgs {
  CtlS* ctl;
  const double* T;
  long long ld;
  int mloc, n, row0, row1;
  double* Acols;
  long long apitch;
  double eps, inf;
  PeerCand* partials;              // [gridDim.x * kSlotStride] ratio-test partials, one 128-byte line each
  unsigned long long* mins;        // [gridDim.x * kMinStride] first improving column of the CTA's range
  unsigned int* syncw;             // words on their own 128-byte lines: ticket A, go A, ticket B, go B, go W
  PeerCand* gwin;                  // sharded: the cross-rank winner, published by CTA 0 for the other CTAs
  long long ll_off;                // sharded: byte offset, inside every rank's exchange block, of the packet
                                   // area: LLPacket row[2][ld], then LLPacket cand[2][kMaxRanks][4]
  Peers peers;
  int rank, world;
  int2* plog;
  long long log_cap;
  int* pos2var;
  int block;                       // pending pivots at which the launch stops (the pass comes next)
  unsigned int tag0;               // launch-unique tag base: pivot s of the launch uses tag0 + s in both of its syncs
};

__device__ __forceinline__ double ldcg_f64(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ldcg_s32(const int* p) {
  int v;
  asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_relaxed_gpu_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr unsigned long long kPanelSpinNs = 20ull * 1000ull * 1000ull * 1000ull;

#ifdef LPS_PANEL_TIMING
// phase clock of CTA 0 / thread 0, kept in registers and written out when the kernel leaves
#define PANEL_MARK(k)                                        \
  do {                                                       \
    if (cta == 0 && tid == 0) {                              \
      unsigned long long now_ = globaltimer_ns();            \
      dbg_[k] += now_ - mark_;                               \
      mark_ = now_;                                          \
    }                                                        \
  } while (0)
#define PANEL_DUMP()                                         \
  do {                                                       \
    if (cta == 0 && tid == 0) {                              \
      _Pragma("unroll") for (int q_ = 0; q_ < 12; q_++) ctl->dbg_ns[q_] += dbg_[q_]; \
      ctl->dbg_ns[12] += (unsigned long long)(clock64() - clk0_);                      \
      ctl->dbg_ns[13] += globaltimer_ns() - ns0_;                                      \
    }                                                        \
  } while (0)
#else
#define PANEL_MARK(k) do { } while (0)
#define PANEL_DUMP() do { } while (0)
#endif

// Arrive-and-wait of the whole co-resident grid, shaped for latency: every CTA's thread 0 fences
// and takes a ticket; the last arriver re-arms the ticket counter and stores the step's tag into
// ONE go word that the other CTAs' thread 0 poll (relaxed loads, then one acquire fence).
// Whatever the CTAs wrote before the call (their slots, their share of the pending row / column)
// is visible to every CTA after it.  False if a peer raised ctl->abort or the wait timed out.
__device__ __forceinline__ bool panel_sync(CtlS* ctl, unsigned int* counter, unsigned int* go, unsigned int tag) {
  __shared__ int s_alive;
  __syncthreads();
  if (threadIdx.x == 0) {
    int alive = 1;
    __threadfence();
    const unsigned int old = atomicAdd(counter, 1u);
    if (old == gridDim.x - 1) {
      atomicExch(counter, 0u);
      __threadfence();
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(go), "r"(tag) : "memory");
    } else if (ld_relaxed_gpu_u32(go) != tag) {
      const unsigned long long t0 = globaltimer_ns();
      unsigned int spins = 0;
      while (ld_relaxed_gpu_u32(go) != tag) {
        if ((++spins & 255u) == 0 &&
            (globaltimer_ns() - t0 > kPanelSpinNs || ldcg_s32(&ctl->abort) != 0)) { alive = 0; break; }
      }
    }
    // acquire side of the hand-off: the relaxed polls above carry no ordering, so one gpu-scope fence once
    // the go word has been seen orders every later read of this CTA (after the closing bar.sync) behind the
    // producers' fenced tickets (round 1 relied on L1-bypassing loads alone: a data race under the PTX model)
    fence_acq_rel_gpu();
    s_alive = alive;
  }
  __syncthreads();
  return s_alive != 0;
}

// one out-of-line copy of the division (the panel runs once per pivot out of a cold instruction
// cache: code size is latency here)
__device__ __noinline__ double ddiv_call(double a, double b) { return __ddiv_rn(a, b); }

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  unsigned int sa = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem_src) : "memory");
}

template <bool kSharded>
__global__ void __launch_bounds__(kPanelThreads, 1) kb_panel(const __grid_constant__ PanelArgs a) {
  CtlS* const ctl = a.ctl;
  if (ctl->base.status != kRunning) return;
  extern __shared__ __align__(16) double s_op[];     // [kPanelMax][kPanelThreads] operand staging (cp.async)
  __shared__ double s_p[kPanelMax], s_re[kPanelMax], s_rn[kPanelMax], s_al[kPanelMax], s_am[kPanelMax];
  __shared__ int s_l[kPanelMax], s_e[kPanelMax];
  __shared__ PeerCand s_red[kPanelThreads / 32], s_red2[kPanelThreads / 32];
  __shared__ int s_min[kPanelThreads / 32], s_min2[kPanelThreads / 32];
  __shared__ double s_slack, s_pw, s_ce;
  __shared__ int s_row, s_ok, s_e2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x, G = gridDim.x;
  const long long ld = a.ld;
  const int mloc = a.mloc, n = a.n;
  long long np = ctl->base.npivots;
  const long long limit = ctl->base.pivot_limit;
  int t = ctl->blk_pend[0];
  int e = ctl->e_nx[(np + 1) & 1];
  if (tid < t) {
    s_l[tid] = ctl->blk_l2[0][tid];
    s_e[tid] = ctl->blk_e2[0][tid];
    s_p[tid] = ctl->blk_p2[0][tid];
  }
  // my share of the rows in phase A (objective row included) and of the columns in phase B:
  // contiguous ranges, normally one row / one column per thread
  const int RW = (mloc + 1 + G - 1) / G;
  const int ilo = cta * RW, ihi = min(ilo + RW, mloc + 1);
  const int W = (int)((((ld + G - 1) / G) + 3) / 4 * 4);
  const long long jlo = (long long)cta * W, jhi = (jlo + W < ld) ? jlo + W : ld;
  const double* const rows = a.peers.rowbuf[a.rank];
  const int scribe = G - 1;   // the CTA with the smallest share writes the shared control words
  unsigned int tag = a.tag0;
  __syncthreads();
#ifdef LPS_PANEL_TIMING
  unsigned long long dbg_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long mark_ = globaltimer_ns();
  const long long clk0_ = clock64();
  const unsigned long long ns0_ = mark_;
#endif

  while (t < a.block) {
    const unsigned int seq = (unsigned int)(np + 1);
    const int par = seq & 1;
    // ---------------- phase A: entering column, b column, ratio test ----------------
    PeerCand best;
    best.slack = a.inf; best.row = kNone; best.p = 0.0;
    if (e != kNone) {
      if (tid < t) {     // pending rows' entries in columns e and n (written by other CTAs: past the gather)
        s_rn[tid] = ldcg_f64(rows + (long long)tid * ld + n);
        s_re[tid] = ldcg_f64(rows + (long long)tid * ld + e);
      }
      double* acol = a.Acols + (long long)t * a.apitch;
      for (int i0 = ilo; i0 < ihi; i0 += kPanelThreads) {   // one trip unless m + 1 > threads * gridDim.x
        const int i = i0 + tid;
        double xe = 0.0, xb = 0.0;
        if (i < ihi) {
          for (int u = 0; u < t; u++)                       // pending columns' entries of my row: all in flight at once
            cp_async8(s_op + u * kPanelThreads + tid, a.Acols + (long long)u * a.apitch + i);
          xe = a.T[(long long)i * ld + e];
          xb = a.T[(long long)i * ld + n];
        }
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();   // s_re / s_rn
        PANEL_MARK(6);
        if (i < ihi) {
          for (int u = 0; u < t; u++) {
            const double au = s_op[u * kPanelThreads + tid];
            if (i == s_l[u]) {                              // row i is pending pivot u's scaled leaving row
              xe = s_re[u];
              xb = s_rn[u];
            } else {
              xe = (e == s_e[u]) ? -ddiv_call(au, s_p[u]) : __dsub_rn(xe, __dmul_rn(au, s_re[u]));
              xb = __dsub_rn(xb, __dmul_rn(au, s_rn[u]));
            }
          }
          acol[i] = xe;
          if (i < mloc && !(xe < a.eps)) {                 // LPState.java:294-299
            double sl = ddiv_call(xb, xe);
            if (sl < best.slack) { best.slack = sl; best.row = i; best.p = xe; }
          }
        }
        __syncthreads();   // s_op is reused by the next trip / phase B
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      double os = __shfl_xor_sync(0xffffffffu, best.slack, off);
      double op = __shfl_xor_sync(0xffffffffu, best.p, off);
      int orow = __shfl_xor_sync(0xffffffffu, best.row, off);
      if (os < best.slack || (os == best.slack && orow < best.row)) { best.slack = os; best.p = op; best.row = orow; }
    }
    PANEL_MARK(7);
    if (lane == 0) s_red[warp] = best;
    __syncthreads();
    PANEL_MARK(8);
    if (tid == 0) {
      PeerCand c = s_red[0];
      for (int w = 1; w < kPanelThreads / 32; w++) {
        const PeerCand o = s_red[w];
        if (o.slack < c.slack || (o.slack == c.slack && o.row < c.row)) c = o;
      }
      PeerCand* mine = &a.partials[cta * kSlotStride];
      mine->slack = c.slack;
      mine->p = c.p;
      mine->row = c.row;
    }
    PANEL_MARK(0);
    // all-gather of the partials: grid-wide arrive-and-wait, then thread k takes CTA k's slot
    if (!panel_sync(ctl, a.syncw + 0, a.syncw + 32, tag)) {
      if (tid == 0) {
        ctl->base.status = kCommTimeout;
        ctl->abort = 1;
        __threadfence();
      }
      return;
    }
    {
      PeerCand c;
      c.slack = a.inf; c.row = kNone; c.p = 0.0;
      for (int k = tid; k < G; k += kPanelThreads) {
        const PeerCand* src = &a.partials[k * kSlotStride];
        PeerCand o;
        o.slack = ldcg_f64(&src->slack);
        o.p = ldcg_f64(&src->p);
        o.row = ldcg_s32(&src->row);
        if (o.slack < c.slack || (o.slack == c.slack && o.row < c.row)) c = o;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        double os = __shfl_xor_sync(0xffffffffu, c.slack, off);
        double op = __shfl_xor_sync(0xffffffffu, c.p, off);
        int orow = __shfl_xor_sync(0xffffffffu, c.row, off);
        if (os < c.slack || (os == c.slack && orow < c.row)) { c.slack = os; c.p = op; c.row = orow; }
      }
      if (lane == 0) s_red2[warp] = c;
      __syncthreads();
    }
    PANEL_MARK(1);
    if (warp == 0) {
      PeerCand c;
      c.slack = a.inf; c.row = kNone; c.p = 0.0;
      if (lane < kPanelThreads / 32) c = s_red2[lane];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        double os = __shfl_xor_sync(0xffffffffu, c.slack, off);
        double op = __shfl_xor_sync(0xffffffffu, c.p, off);
        int orow = __shfl_xor_sync(0xffffffffu, c.row, off);
        if (os < c.slack || (os == c.slack && orow < c.row)) { c.slack = os; c.p = op; c.row = orow; }
      }
      int ok = 1;
      if (kSharded) {
        // only CTA 0 talks to the peers (148 CTAs polling the same two mailbox lines would fight the
        // incoming NVLink writes); it hands the cross-rank winner to the other CTAs through one go word
        if (cta == 0) {
          if (lane < a.world) {   // one lane per peer (own packet slots included): three packets, no fence
            LLPacket* dst = reinterpret_cast<LLPacket*>(reinterpret_cast<char*>(a.peers.blk[lane]) + a.ll_off) +
                            2 * ld + ((size_t)par * kMaxRanks + a.rank) * 4;
            ll_store(dst + 0, c.slack, seq);
            ll_store(dst + 1, c.p, seq);
            ll_store(dst + 2, (double)((c.row == kNone) ? -1 : a.row0 + c.row), seq);   // row < 2^31: exact
          }
          PeerCand pc;
          pc.slack = a.inf; pc.row = kNone; pc.p = 0.0;
          if (lane < a.world) {
            const LLPacket* src = reinterpret_cast<const LLPacket*>(reinterpret_cast<const char*>(a.peers.blk[a.rank]) + a.ll_off) +
                                  2 * ld + ((size_t)par * kMaxRanks + lane) * 4;
            double rowd = -1.0;
            ok = (ll_load(src + 0, seq, pc.slack) && ll_load(src + 1, seq, pc.p) && ll_load(src + 2, seq, rowd)) ? 1 : 0;
            pc.row = (rowd < 0.0) ? kNone : (int)rowd;
          }
          ok = __all_sync(0xffffffffu, ok) ? 1 : 0;
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            double os = __shfl_xor_sync(0xffffffffu, pc.slack, off);
            double op = __shfl_xor_sync(0xffffffffu, pc.p, off);
            int orow = __shfl_xor_sync(0xffffffffu, pc.row, off);
            if (os < pc.slack || (os == pc.slack && orow < pc.row)) { pc.slack = os; pc.p = op; pc.row = orow; }
          }
          c = pc;   // GLOBAL row from here on
          if (lane == 0) {
            if (ok) {
              a.gwin->slack = c.slack;
              a.gwin->p = c.p;
              a.gwin->row = c.row;
              st_release_gpu_u32(a.syncw + 128, tag);
            } else {
              ctl->base.status = kCommTimeout;
              ctl->abort = 1;
              __threadfence();
            }
          }
        } else {
          if (lane == 0) {
            if (ld_relaxed_gpu_u32(a.syncw + 128) != tag) {
              const unsigned long long t0 = globaltimer_ns();
              unsigned int spins = 0;
              while (ld_relaxed_gpu_u32(a.syncw + 128) != tag) {
                if ((++spins & 255u) == 0 &&
                    (globaltimer_ns() - t0 > kPanelSpinNs || ldcg_s32(&ctl->abort) != 0)) { ok = 0; break; }
              }
            }
            fence_acq_rel_gpu();      // acquire: pairs with CTA 0's st.release of the go word
            c.slack = ldcg_f64(&a.gwin->slack);
            c.p = ldcg_f64(&a.gwin->p);
            c.row = ldcg_s32(&a.gwin->row);
          }
        }
      }
      if (lane == 0) { s_slack = c.slack; s_pw = c.p; s_row = c.row; s_ok = ok; }
    }
    __syncthreads();
    const int l = (s_row == kNone) ? -1 : s_row;       // global row (== local on a single GPU)
    const double p = s_pw;
    PANEL_MARK(2);
    if (kSharded && !s_ok) {        // a peer never delivered its candidate: everybody out
      if (tid == 0) {
        ctl->base.status = kCommTimeout;
        ctl->abort = 1;
        __threadfence();
      }
      return;
    }
    int verdict = kRunning;
    if (e == kNone) verdict = kOptimal;                 // getEntering() == -1   LPSolver.java:101
    else if (l < 0) verdict = kUnbounded;               // getLeaving() == -1    LPSolver.java:103
    else if (np >= limit) verdict = kPivotCap;
    if (verdict != kRunning) {
      if (cta == scribe && tid == 0) {
        ctl->base.status = verdict;
        ctl->base.e_cur = (e == kNone) ? -1 : e;
        ctl->base.l_cur = (verdict == kPivotCap) ? l : -1;
        ctl->e_nx[par] = e;
      }
      PANEL_DUMP();
      return;
    }
    // ---------------- phase B: leaving row, objective row, next entering column ----------------
    const bool i_own = !kSharded || (l >= a.row0 && l < a.row1);
    const int lloc = i_own ? l - a.row0 : -1;
    if (tid < t) {
      s_al[tid] = i_own ? ldcg_f64(a.Acols + (long long)tid * a.apitch + lloc) : 0.0;
      s_am[tid] = ldcg_f64(a.Acols + (long long)tid * a.apitch + mloc);
    }
    if (tid == kPanelThreads - 1) s_ce = ldcg_f64(a.Acols + (long long)t * a.apitch + mloc);
    int mine_next = kNone;
    for (long long jb = jlo; jb < jhi; jb += kPanelThreads) {   // one trip unless ld > threads * gridDim.x
      const long long j = jb + tid;
      double xc = 0.0, x = 0.0;
      if (j < jhi) {
        for (int u = 0; u < t; u++)                         // pending rows' entries of my column
          cp_async8(s_op + u * kPanelThreads + tid, rows + (long long)u * ld + j);
        if (j < n) xc = a.T[(long long)mloc * ld + j];
        if (i_own && j <= n) x = a.T[(long long)lloc * ld + j];
      }
      cp_async_commit();
      cp_async_wait_all();
      __syncthreads();     // s_al / s_am / s_ce
      if (jb == jlo) PANEL_MARK(9);
      bool got = true;
      if (j < jhi) {
        const double ce = s_ce;
        for (int u = 0; u < t; u++) {
          const double ru = s_op[u * kPanelThreads + tid];
          const bool pcol = ((int)j == s_e[u]);
          const double pu = s_p[u];
          if (lloc >= 0 && lloc == s_l[u]) x = ru;          // the row was pending pivot u's leaving row
          else x = pcol ? -ddiv_call(s_al[u], pu) : __dsub_rn(x, __dmul_rn(s_al[u], ru));
          xc = pcol ? -ddiv_call(s_am[u], pu) : __dsub_rn(xc, __dmul_rn(s_am[u], ru));
        }
        double r = 0.0;
        if (i_own) {
          if (j <= n) r = ((int)j == e) ? ddiv_call(1.0, p) : ddiv_call(x, p);      // LPState.java:139-146
          if (kSharded) {      // compute + broadcast in one kernel: one packet per peer, straight into its memory
            for (int k = 0; k < a.world; k++)
              if (k != a.rank)
                ll_store(reinterpret_cast<LLPacket*>(reinterpret_cast<char*>(a.peers.blk[k]) + a.ll_off) +
                             (size_t)par * ld + j, r, seq);
          }
        } else {
          got = ll_load(reinterpret_cast<const LLPacket*>(reinterpret_cast<const char*>(a.peers.blk[a.rank]) + a.ll_off) +
                            (size_t)par * ld + j, seq, r);
        }
        (a.peers.rowbuf[a.rank] + (long long)t * ld)[j] = r;     // this rank's copy of the pending row
        if (j < n) {
          double cn = ((int)j == e) ? -ddiv_call(ce, p) : __dsub_rn(xc, __dmul_rn(ce, r));   // :170-178
          if (cn > a.eps && (int)j < mine_next) mine_next = (int)j;
        }
      }
      if (__syncthreads_or(got ? 0 : 1)) {   // (also: s_op is reused by the next trip / phase A)
        if (tid == 0) {        // the owner's packets never came
          ctl->base.status = kCommTimeout;
          ctl->abort = 1;
          __threadfence();
        }
        return;
      }
    }
    PANEL_MARK(10);
    mine_next = warp_min_int(mine_next);
    if (lane == 0) s_min[warp] = mine_next;
    if (tid == 0) {          // every CTA keeps its own copy of the pending pivots' scalars
      s_e[t] = e;
      s_l[t] = lloc;
      s_p[t] = p;
    }
    __syncthreads();
    PANEL_MARK(11);
    if (tid == 0) {
      int v = s_min[0];
      for (int w = 1; w < kPanelThreads / 32; w++) v = min(v, s_min[w]);
      a.mins[cta * kMinStride] = (unsigned long long)(unsigned int)v;
    }
    PANEL_MARK(3);
    // all-gather of the per-CTA minima -> the next entering column (LPState.java:274-285)
    if (!panel_sync(ctl, a.syncw + 64, a.syncw + 96, tag)) {
      if (tid == 0) {
        ctl->base.status = kCommTimeout;
        ctl->abort = 1;
        __threadfence();
      }
      return;
    }
    {
      int v = kNone;
      for (int k = tid; k < G; k += kPanelThreads) {
        unsigned long long w;
        asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(w) : "l"(a.mins + k * kMinStride) : "memory");
        v = min(v, (int)(unsigned int)(w & 0xffffffffull));
      }
      v = warp_min_int(v);
      if (lane == 0) s_min2[warp] = v;
      __syncthreads();
      if (tid == 0) {
        int m2 = s_min2[0];
        for (int w = 1; w < kPanelThreads / 32; w++) m2 = min(m2, s_min2[w]);
        s_e2 = m2;
      }
      __syncthreads();
    }
    PANEL_MARK(4);
    const int e2 = s_e2;
    if (cta == scribe && tid == 0) {                       // commit pivot(e, l)
      ctl->base.e_cur = e;
      ctl->base.l_cur = l;
      ctl->base.p = p;
      ctl->owner = i_own ? a.rank : -1;
      ctl->blk_e2[0][t] = e;
      ctl->blk_l2[0][t] = lloc;
      ctl->blk_p2[0][t] = p;
      ctl->blk_pend[0] = t + 1;
      ctl->e_nx[par ^ 1] = e2;
      a.plog[np % a.log_cap] = make_int2(e, l);
      int tmp = a.pos2var[e];                              // exchangeIndexes, LPState.java:311-320
      a.pos2var[e] = a.pos2var[n + l];
      a.pos2var[n + l] = tmp;
      ctl->base.npivots = np + 1;
    }
    np += 1;
    t += 1;
    e = e2;
    tag += 1;
    PANEL_MARK(5);
  }
  PANEL_DUMP();
}

// K3 (blocked): apply all pending pivots to every local cell (objective row included) in one
// pass: one 256-bit load and one 256-bit store per four cells, 2t flops per cell in between.
//
// One persistent CTA per SM, kLanes x 128 threads.  Work is cut into chunks of kCH rows x 512
// columns, numbered row-band-major (consecutive chunks are neighbouring strips of the same rows, so
// the CTAs of the grid sweep the tableau as one band: DRAM pages stay open) and claimed from an
// atomic counter one chunk ahead.  Each 128-thread lane owns whole groups of kU rows (kU 256-bit
// loads in flight per thread).  The operands of the replay are double-buffered in shared memory
// and arrive by cp.async while the previous chunk is being computed:
//   s_r [2][t][512]   the strip's slice of the pending rows
//   s_a [2][t][kCH]   the chunk's slice of the pending columns
// so the inner loop is LDS + DMUL/DADD only, branch-free for every cell; the few cells a pending
// pivot overwrites (its leaving row, its entering column) are recomputed afterwards.
// A subset of a CTA's threads that works as a unit: its own thread numbering and its own hardware barrier.
// kBarId 0 = the whole CTA (__syncthreads); otherwise threads [kTid0, kTid0 + kCount) meet on named barrier kBarId,
// so that two groups of one CTA (the pass warps and the panel warps of kb_step_ws) never wait for each other.
template <int kBarId, int kCount, int kTid0>
struct CtaGroup {
  static __device__ __forceinline__ int tid() { return (int)threadIdx.x - kTid0; }
  static __device__ __forceinline__ void sync() {
    if (kBarId == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"n"(kBarId), "n"(kCount) : "memory");
  }
  static __device__ __forceinline__ int sync_or(int pred) {
    if (kBarId == 0) return __syncthreads_or(pred);
    int r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.s32 p, %3, 0;\n\t"
        "barrier.cta.red.or.pred q, %1, %2, p;\n\t"
        "selp.s32 %0, 1, 0, q;\n\t}"
        : "=r"(r)
        : "n"(kBarId), "n"(kCount), "r"(pred)
        : "memory");
    return r;
  }
};
using WholeCta = CtaGroup<0, 0, 0>;

constexpr int kFlushThreads = 128;
constexpr int kStripCols = 4 * kFlushThreads;


// The pass as a role: CTAs that call it (ncta of them) apply pending set `set` to the tableau Tsrc -> Tdst
// (the same buffer for an in-place pass).  Acols / Rrows: the set's pending columns / rows.  `smem`: the CTA's
// dynamic shared memory, 2 t (512 + kCH) doubles.  The last CTA retires the set (see sweep_role in lps_sweep.cuh:
// same protocol, so the two pass kernels are interchangeable inside kb_step).
struct FlushArgs {
  CtlS* ctl;
  double* Tbuf[2];
  long long ld;
  int mloc;
  const double* Acols;     // [2][block][apitch]
  long long apitch;
  const double* Rrows;     // [2][block][ld]  (this rank's copy)
  int block;
  int q;                   // launch parity = pending set
  int inplace;
  int ncta;
  int hints;               // bit 0: evict_first on the stream's loads, bit 1: on its stores
};

template <int kLanes, int kU, int kG, bool kPre, class Grp = WholeCta>   // kPre: prefetch the lane's next group of rows into L2
__device__ __forceinline__ void flush_role(const FlushArgs& fa, double* smem) {
  constexpr int kThreads = kFlushThreads * kLanes;
  constexpr int kCH = kU * kLanes * kG;            // rows per chunk
  static_assert(kU % 2 == 0, "16-byte copies of the pending columns");
  CtlS* const ctl = fa.ctl;
  const int set = fa.q;
  const int t = ctl->blk_pend[set];
  const int cur = ctl->cur_at[fa.q];
  const double* const Ts = fa.Tbuf[cur];
  double* const T = fa.Tbuf[fa.inplace ? cur : (cur ^ 1)];
  const long long ld = fa.ld;
  const int mloc = fa.mloc;
  const long long apitch = fa.apitch;
  const double* const Acols = fa.Acols + (size_t)set * fa.block * apitch;
  const double* const Rrows = fa.Rrows + (size_t)set * fa.block * ld;
  __shared__ bool s_last;
  __shared__ unsigned long long s_t0;
  if (Grp::tid() == 0) s_t0 = globaltimer_ns();
  // out of place = beside a panel (look-ahead loop): mark the stream evict_first so that the panel's operands
  // survive in L2; the stand-alone in-place pass keeps the plain accesses it was tuned with
  const bool kStreamL = !fa.inplace && (fa.hints & 1), kStreamS = !fa.inplace && (fa.hints & 2);
  const unsigned long long pol_stream = l2_evict_first_policy();
  if (t > 0) {
  double* const s_r = smem;                                      // [2][t][kStripCols]
  double* const s_a = smem + (size_t)2 * t * kStripCols;         // [2][t][kCH]
  __shared__ double s_p[kMaxBlock];
  __shared__ int s_l[kMaxBlock], s_e[kMaxBlock];
  __shared__ long long s_claim[2];
  const int tid = Grp::tid(), lane = tid / kFlushThreads, ltid = tid % kFlushThreads;
  if (tid < t) {
    s_l[tid] = ctl->blk_l2[set][tid];
    s_e[tid] = ctl->blk_e2[set][tid];
    s_p[tid] = ctl->blk_p2[set][tid];
  }
  const int nstrips = (int)((ld + kStripCols - 1) / kStripCols);
  const int nrb = (mloc + 1 + kCH - 1) / kCH;                    // chunks per strip
  const long long nchunks = (long long)nstrips * nrb;
  unsigned long long* const queue = &ctl->blk_queue;
  // a CTA holds the chunk it is replaying and ONE more (whose operand slices are in flight): claiming
  // further ahead leaves the last CTAs with a private backlog while the others idle, which at a few
  // chunks per CTA (8-way shards, 10,000 x 10,000) was more than half of the kernel
  if (tid == 0) s_claim[0] = (long long)atomicAdd(queue, 1ull);
  Grp::sync();
  long long cur = s_claim[0], nxt = 0;

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
  while (cur < nchunks) {
    if (tid == 0) s_claim[buf ^ 1] = (long long)atomicAdd(queue, 1ull);   // the chunk after this one
    cp_async_wait_all();
    Grp::sync();   // this chunk's operands have landed; everyone is done with the other buffer
    nxt = s_claim[buf ^ 1];
    if (nxt < nchunks) prefetch(nxt, buf ^ 1);

    const long long j0 = (cur % nstrips) * kStripCols + 4 * ltid;
    {
      const bool active = (j0 < ld);     // the last strip can be narrower than the CTA
      const int i0 = (int)(cur / nstrips) * kCH;
      const int i_end = min(i0 + kCH, mloc + 1);
      const double* sr = s_r + (size_t)buf * t * kStripCols + 2 * ltid;
      const double* sa = s_a + (size_t)buf * t * kCH;
      // Cells that a pending pivot OVERWRITES rather than updates are rare: a thread's four columns hold
      // one of the t entering columns (cmask, per thread), or the chunk holds one of the t leaving rows
      // (rmask, same for the whole CTA; only the last pivot on a row counts).  Everybody first runs the
      // branch-free replay on every cell, then those cells are recomputed from the pivot that overwrote
      // them — the same values, since nothing before an overwrite survives it — so a warp with a special
      // column or row costs a few dozen instructions more instead of taking a branchy path for the
      // whole group (which left its lane late at every chunk barrier).
      unsigned int cmask = 0, rmask = 0;
      for (int u = 0; u < t; u++) {
        if (s_e[u] >= j0 && s_e[u] < j0 + 4) cmask |= 1u << u;
        if (s_l[u] >= i0 && s_l[u] < i_end) {
          bool last = true;
          for (int w = u + 1; w < t; w++) last &= (s_l[w] != s_l[u]);
          if (last) rmask |= 1u << u;
        }
      }
      double* base = T + j0;               // stores
      const double* sbase = Ts + j0;       // loads (the same buffer unless the pass runs out of place)
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
          if (i + k < i_end) x[k] = kStreamL ? ld256_hint(sbase + (long long)(i + k) * ld, pol_stream) : ld256(sbase + (long long)(i + k) * ld);
          else x[k].x = x[k].y = x[k].z = x[k].w = 0.0;
      };
      // replay the pending pivots on one group of rows held in registers, then store it
      auto finish_group = [&](D4 (&x)[kU], int i, int g) {
        // 1. every pending pivot on every cell, branch-free; the operands of pivot u + 1 are fetched from
        //    shared memory before the 8 kU FP64 instructions of pivot u are issued
        {
          auto load_ops = [&](int u, D4& r, double (&av)[kU]) {
            r = rvec(u);
#pragma unroll
            for (int q = 0; q < kU / 2; q++) {
              const double2 v = *reinterpret_cast<const double2*>(sa + u * kCH + g * kU + 2 * q);
              av[2 * q] = v.x;
              av[2 * q + 1] = v.y;
            }
          };
          auto apply = [&](const D4& r, const double (&av)[kU]) {
#pragma unroll
            for (int k = 0; k < kU; k++) {
              x[k].x = __dsub_rn(x[k].x, __dmul_rn(av[k], r.x));
              x[k].y = __dsub_rn(x[k].y, __dmul_rn(av[k], r.y));
              x[k].z = __dsub_rn(x[k].z, __dmul_rn(av[k], r.z));
              x[k].w = __dsub_rn(x[k].w, __dmul_rn(av[k], r.w));
            }
          };
          D4 r0, r1;
          double a0[kU], a1[kU];
          load_ops(0, r0, a0);
          int u = 0;
          for (; u + 2 <= t; u += 2) {
            load_ops(u + 1, r1, a1);
            apply(r0, a0);
            if (u + 2 < t) load_ops(u + 2, r0, a0);
            apply(r1, a1);
          }
          if (u < t) apply(r0, a0);
        }
        // 2. my columns that were an entering column: -(a/p) at the last such pivot, plain updates after it
        if (cmask != 0) {
#pragma unroll
          for (int c = 0; c < 4; c++) {      // c is a compile-time constant: x stays in registers
            int us = -1;
            for (int u = 0; u < t; u++)
              if (((cmask >> u) & 1u) && s_e[u] - j0 == c) us = u;
            if (us < 0) continue;
#pragma unroll
            for (int k = 0; k < kU; k++) {
              double y = -__ddiv_rn(sa[us * kCH + g * kU + k], s_p[us]);        // LPState.java:157 / :172
              for (int u = us + 1; u < t; u++) {
                const D4 r = rvec(u);
                const double rc = (c == 0) ? r.x : (c == 1) ? r.y : (c == 2) ? r.z : r.w;
                y = __dsub_rn(y, __dmul_rn(sa[u * kCH + g * kU + k], rc));
              }
              if (c == 0) x[k].x = y; else if (c == 1) x[k].y = y; else if (c == 2) x[k].z = y; else x[k].w = y;
            }
          }
        }
        // 3. rows of this group that were a leaving row: the scaled row of the last such pivot, then the
        //    later pivots on it (a later entering column among my four overwrites again)
        if (rmask != 0) {
          for (int us = 0; us < t; us++) {
            if (!((rmask >> us) & 1u) || s_l[us] < i || s_l[us] >= i + kU) continue;
            const int ks = s_l[us] - i;
            D4 y = rvec(us);                                                         // LPState.java:137-146
            for (int u = us + 1; u < t; u++) {
              const D4 r = rvec(u);
              const double a = sa[u * kCH + g * kU + ks];
              const int ce = ((cmask >> u) & 1u) ? (int)(s_e[u] - j0) : -1;
              const double q = (ce >= 0) ? -__ddiv_rn(a, s_p[u]) : 0.0;
              y.x = (ce == 0) ? q : __dsub_rn(y.x, __dmul_rn(a, r.x));
              y.y = (ce == 1) ? q : __dsub_rn(y.y, __dmul_rn(a, r.y));
              y.z = (ce == 2) ? q : __dsub_rn(y.z, __dmul_rn(a, r.z));
              y.w = (ce == 3) ? q : __dsub_rn(y.w, __dmul_rn(a, r.w));
            }
#pragma unroll
            for (int k = 0; k < kU; k++)
              if (k == ks) x[k] = y;
          }
        }
#pragma unroll
        for (int k = 0; k < kU; k++)
          if (i + k < i_end) {
            if (kStreamS) st256_hint(base + (long long)(i + k) * ld, x[k], pol_stream);
            else st256(base + (long long)(i + k) * ld, x[k]);
          }
      };
      // the lane's NEXT group of rows (in this chunk, else the first one of the next chunk) is pulled
      // into L2 while this group is replayed: prefetches hold no register and no scoreboard
      auto prefetch_rows = [&](const double* b, int i, int iend) {
        if ((ltid & 3) == 0) {             // one 128-byte line per four threads
#pragma unroll
          for (int k = 0; k < kU; k++)
            if (i + k < iend) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + (long long)(i + k) * ld));
        }
      };
      // (claiming groups dynamically inside a chunk was tried: the two lane barriers per group cost
      // more than the imbalance they remove — profiles/r01_flush_variants.md)
      if (active) {
        for (int g = lane; g < kCH / kU; g += kLanes) {
          const int i = i0 + g * kU;
          if (i >= i_end) break;
          D4 x[kU];
          load_group(x, i);
          if (kPre) {
            const int in = i + kLanes * kU;
            if (g + kLanes < kCH / kU && in < i_end) {
              prefetch_rows(sbase, in, i_end);
            } else if (nxt < nchunks) {
              const long long j0n = (nxt % nstrips) * kStripCols + 4 * ltid;
              const int i0n = (int)(nxt / nstrips) * kCH;
              if (j0n < ld) prefetch_rows(Ts + j0n, i0n + lane * kU, min(i0n + kCH, mloc + 1));
            }
          }
          finish_group(x, i, g);
        }
      }
    }
    cur = nxt;
    buf ^= 1;
  }
  }
  // last pass CTA retires the block
  Grp::sync();
  if (Grp::tid() == 0) {
    __threadfence();
    unsigned int tk = atomicAdd(&ctl->blk_ticket, 1u);
    s_last = (tk == (unsigned int)fa.ncta - 1);
  }
  Grp::sync();
  if (s_last && Grp::tid() == 0) {
    ctl->blk_ticket = 0;
    ctl->blk_queue = 0;
    ctl->blk_pend[set] = 0;
    ctl->cur_at[fa.q ^ 1] = (!fa.inplace && t > 0) ? (cur ^ 1) : cur;   // read by the NEXT launch only
    if (t > 0) {
      ctl->sweeps_done += 1;
      ctl->dbg_ns[10] += globaltimer_ns() - s_t0;      // the pass's own clock: the last CTA to retire (host: split tuning)
      ctl->dbg_ns[11] += 1;
    }
    __threadfence();
  }
}

// the pass as a kernel of its own, in place on set 0 (loop modes 5 and 6)
template <int kLanes, int kU, int kG, bool kPre>
__global__ void __launch_bounds__(kFlushThreads * kLanes, 1)
kb_flush(CtlS* ctl, double* T, long long ld, int mloc, const double* Acols, long long apitch, const double* Rrows,
         int block) {
  extern __shared__ __align__(32) double flush_smem[];
  FlushArgs fa;
  fa.ctl = ctl;
  fa.Tbuf[0] = fa.Tbuf[1] = T;
  fa.ld = ld;
  fa.mloc = mloc;
  fa.Acols = Acols;
  fa.apitch = apitch;
  fa.Rrows = Rrows;
  fa.block = block;
  fa.q = 0;
  fa.inplace = 1;
  fa.hints = 0;
  fa.ncta = (int)gridDim.x;
  flush_role<kLanes, kU, kG, kPre>(fa, flush_smem);
}



}  // namespace lps
