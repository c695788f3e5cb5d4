// Row-sharded pivot loop: one rank (process or device) per GPU of one NVSwitch box.
//
// Partition (the reference's own thread split, LPState.java:222-223): rank k owns the rows
// [k*m/G, (k+1)*m/G) of (A | b); the objective row (c | -v) is REPLICATED on every rank and
// updated redundantly from the same pivot row, so every rank derives the same entering
// column with no communication.
//
// Exchange per pivot, fused into the kernels and carried by plain stores into PEER memory over
// NVLink (no NCCL call, no host in the loop):
//   ks_ratio      local ratio test; its last block PUSHES the rank's best (ratio, row, pivot
//                 element) into slot[rank] of every peer's mailbox and releases a sequence flag.
//   ks_scale_row  every CTA acquires the G candidates, takes the lexicographic minimum (lowest
//                 global row wins ties — the sequential rule of LPState.java:292-303), and
//                   owner of row l : scales its 256-column chunk of the pivot row and stores it
//                                    into its own and every peer's row buffer, then releases a
//                                    per-chunk flag  (compute + broadcast in ONE kernel);
//                   other ranks    : acquire the chunk's flag and read the chunk locally;
//                 all ranks then evaluate the new objective chunk to find the next entering column.
//   ks_update     the same streaming update as the single-GPU kernel on the local rows.
// Mailboxes and row buffers are double-buffered by pivot parity; a rank can run at most one
// pivot ahead of its slowest peer (it needs that peer's next candidate to go further).
#pragma once
#include "lps_kernels.cuh"

namespace lps {

constexpr int kMaxRanks = 8;
constexpr int kChunk = 256;           // columns per ks_scale_row CTA == flag granularity
constexpr int kMaxChunks = 4096;      // flag slots per parity (n+1 <= 524288 at 128-column chunks)
constexpr unsigned long long kSpinTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;
constexpr int kMaxBlock = 32;         // blocked loop: most pivots deferred between two tableau passes

enum : int { kCommTimeout = 5 };

struct __align__(32) PeerCand {
  double slack;
  double p;
  int row;            // GLOBAL row index, kNone = no candidate
  unsigned int seq;   // written last, with release semantics
  int pad_[2];
};

// One per rank, in that rank's device memory; peers write into it.
struct CommBlock {
  PeerCand cand[2][kMaxRanks];
  unsigned int row_flag[2][kMaxChunks];
  // followed by: double rowbuf[2][ld]
};

struct Peers {
  CommBlock* blk[kMaxRanks];
  double* rowbuf[kMaxRanks];  // base of rowbuf[2][ld] inside each rank's block
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_volatile_s32(const int* p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ unsigned int ld_relaxed_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// spin until *flag == want; false on timeout (a dead peer must not hang the GPU).  The polls are
// relaxed loads (an acquire load per poll costs an L1 invalidation each time); one acquire fence
// when the flag has arrived orders everything read afterwards.
template <bool kAcquireFence = true>
__device__ __forceinline__ bool spin_until(const unsigned int* flag, unsigned int want) {
  bool ok = true;
  if (ld_relaxed_sys(flag) != want) {
    const unsigned long long t0 = globaltimer_ns();
    for (;;) {
      bool hit = false;
      for (int k = 0; k < 64 && !hit; k++) hit = (ld_relaxed_sys(flag) == want);
      if (hit) break;
      if (globaltimer_ns() - t0 > kSpinTimeoutNs) { ok = false; break; }
    }
  }
  // callers that go on to read only with L1-bypassing loads (ld.volatile / ld.cg) may skip the fence:
  // the producer fenced before the flag, so the data is in this GPU's memory by the time the flag is
  if (kAcquireFence) fence_acq_rel_sys();
  return ok;
}

// Sharded control fields live in the same Ctl (e_next is unused here):
//   e_nx[seq&1]      entering column of pivot number seq (1-based), atomicMin target of pivot seq-1
struct CtlS {
  Ctl base;
  int e_nx[2];
  int owner;          // rank that owns the leaving row of the pivot in flight
  unsigned int ticket2;
  int abort;          // persistent loop: a CTA gave up (peer timeout); everyone leaves
  int tile_ctr;       // persistent loop: phase-C tile queue
  unsigned long long bar;     // persistent loop: grid-barrier counter
  unsigned long long upd_ns;  // persistent loop: accumulated phase-C time
  // blocked loop (lps_blocked.cuh): pivots committed but not yet applied to the tableau
  int blk_pend[2];            // pending pivots of set 0 / 1, 0..kMaxBlock (the look-ahead loop alternates the sets;
                              // the other blocked loops use set 0 only)
  unsigned int blk_ticket;    // last-CTA-done counter of kb_flush
  unsigned long long blk_queue;   // kb_flush: next unclaimed chunk
  unsigned long long bar_base;    // kb_panel: value of `bar` when the next cooperative launch starts
  // role clocks (ns / counts), cleared by ks_begin_run.  Look-ahead step (lps_step.cuh): [0..5] the panel's phases,
  // [10] pass durations, [11] passes, [12] / [13] the first launch's panel duration / pivots, [14] / [15] panel
  // duration / pivots of the launches with a pass — read by the host to tune the SM split (tune_split, lps_api.cu)
  // and printed with LPS_DEBUG=1.  kb_panel (LPS_PANEL_TIMING builds only): [0..13] phase clock of CTA 0.
  unsigned long long dbg_ns[16];
  int blk_e2[2][kMaxBlock];   // [set][u] entering column of pending pivot u
  int blk_l2[2][kMaxBlock];   // its leaving row as a LOCAL row index, -1 if another rank owns the row
  double blk_p2[2][kMaxBlock];   // its pivot element
  // look-ahead loop (lps_step.cuh): the tableau ping-pongs between two buffers
  int cur_at[2];              // [launch parity] which buffer holds the tableau when that launch starts
  unsigned int sweeps_done;   // passes that applied at least one pivot (host bookkeeping of the timed launches)
  int pad2_;
  int blk_fill[2];            // pivots the panel put into set 0 / 1 (blk_pend is cleared by the pass that applies the
                              // set — possibly while the panel of the same launch is still starting; blk_fill is not)
};

__global__ void ks_begin_run(CtlS* ctl, long long max_pivots, int reset_next) {
  ctl->base.status = kRunning;
  ctl->base.pivot_limit = (max_pivots < 0) ? LLONG_MAX : ctl->base.npivots + max_pivots;
  ctl->base.ticket = 0;
  ctl->ticket2 = 0;
  ctl->abort = 0;
  ctl->bar = 0;
  ctl->bar_base = 0;
  ctl->upd_ns = 0;
  ctl->cur_at[0] = ctl->cur_at[1] = 0;   // the tableau is in the handle's current buffer when a run starts
  ctl->sweeps_done = 0;
  ctl->blk_fill[0] = ctl->blk_fill[1] = 0;
  for (int k = 0; k < 16; k++) ctl->dbg_ns[k] = 0;
  if (reset_next) ctl->e_nx[(ctl->base.npivots + 1) & 1] = kNone;
}

__global__ void ks_first_positive(CtlS* ctl, const double* __restrict__ crow, int n, double eps) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int mine = (j < n && crow[j] > eps) ? j : kNone;
  mine = warp_min_int(mine);
  if ((threadIdx.x & 31) == 0 && mine != kNone) atomicMin(&ctl->e_nx[(ctl->base.npivots + 1) & 1], mine);
}

__global__ void ks_extract(const CtlS* ctl, const double* __restrict__ T, long long ld, int mloc, int n,
                           int e_arg, double* col0, double* col1, double* bcol) {
  int e = (e_arg >= 0) ? e_arg : ctl->e_nx[(ctl->base.npivots + 1) & 1];
  double* col = (ctl->base.npivots & 1) ? col1 : col0;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > mloc) return;
  if (e != kNone) col[i] = T[(long long)i * ld + e];
  bcol[i] = T[(long long)i * ld + n];
}

// K1 (sharded): local ratio test over the rank's rows, then push the candidate to every peer.
__global__ void ks_ratio(CtlS* ctl, const double* __restrict__ col0, const double* __restrict__ col1,
                         const double* __restrict__ bcol, int mloc, int row0, double eps, double inf,
                         Cand* partials, Peers peers, int rank, int world) {
  if (ctl->base.status != kRunning) return;
  const long long np = ctl->base.npivots;
  const unsigned int seq = (unsigned int)(np + 1);
  const double* col = (np & 1) ? col1 : col0;
  Cand best;
  best.slack = inf; best.row = kNone; best.pad_ = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < mloc; i += gridDim.x * blockDim.x) {
    double a = col[i];
    if (!(a < eps)) {
      double s = __ddiv_rn(bcol[i], a);
      if (s < best.slack) { best.slack = s; best.row = i; }
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
      unsigned int t = atomicAdd(&ctl->base.ticket, 1u);
      is_last = (t == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
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
    d.slack = __shfl_sync(0xffffffffu, d.slack, 0);
    d.row = __shfl_sync(0xffffffffu, d.row, 0);
    if (lane == 0) {
      ctl->base.ticket = 0;
      ctl->e_nx[(seq & 1) ^ 1] = kNone;  // slot of pivot seq+1: free since ks_update(seq-1) finished
    }
    // lanes 0..world-1 each push to one peer (own mailbox included)
    if (lane < world) {
      PeerCand* dst = &peers.blk[lane]->cand[seq & 1][rank];
      dst->slack = d.slack;
      dst->p = (d.row == kNone) ? 0.0 : col[d.row];
      dst->row = (d.row == kNone) ? kNone : row0 + d.row;
      __threadfence_system();
      st_release_sys(&dst->seq, seq);
    }
  }
}

// K2 (sharded): candidates -> winner; owner scales + broadcasts its chunk; everyone scans the new
// objective chunk for the next entering column; the last CTA commits the pivot.
__global__ void ks_scale_row(CtlS* ctl, double* __restrict__ T, long long ld, int mloc, int n, int row0,
                             int row1, const double* __restrict__ col0, const double* __restrict__ col1,
                             double eps, double inf, Peers peers, int rank, int world, int2* plog,
                             long long log_cap, int* pos2var) {
  if (ctl->base.status != kRunning) return;
  const long long np = ctl->base.npivots;
  const unsigned int seq = (unsigned int)(np + 1);
  const int par = seq & 1;
  CommBlock* mine = peers.blk[rank];
  __shared__ int s_l, s_verdict;
  __shared__ double s_p;
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
      const int e = ctl->e_nx[par];
      int verdict = kRunning;
      if (!ok) verdict = kCommTimeout;
      else if (e == kNone) verdict = kOptimal;
      else if (c.row == kNone) verdict = kUnbounded;
      else if (np >= ctl->base.pivot_limit) verdict = kPivotCap;
      s_verdict = verdict;
      s_l = (c.row == kNone) ? -1 : c.row;
      s_p = c.p;
    }
  }
  __syncthreads();
  const int verdict = s_verdict;
  const int l = s_l;
  const double p = s_p;
  const int e = ctl->e_nx[par];
  const bool i_own = (l >= row0 && l < row1);
  const int j = blockIdx.x * kChunk + threadIdx.x;
  int mine_next = kNone;
  if (verdict == kRunning) {
    double* rb_local = peers.rowbuf[rank] + (long long)par * ld;
    double r = 0.0;
    if (i_own) {
      if (j < ld) {
        if (j <= n) {
          double* tl = T + (long long)(l - row0) * ld;
          r = (j == e) ? __ddiv_rn(1.0, p) : __ddiv_rn(tl[j], p);
          tl[j] = r;
        }
        for (int k = 0; k < world; k++) (peers.rowbuf[k] + (long long)par * ld)[j] = r;
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
      if (j < ld) r = ld_volatile_f64(rb_local + j);
    }
    if (j < n) {
      const double* col = (np & 1) ? col1 : col0;
      const double ce = col[mloc];
      double cj = T[(long long)mloc * ld + j];
      double cn = (j == e) ? -__ddiv_rn(ce, p) : __dsub_rn(cj, __dmul_rn(ce, r));
      if (cn > eps) mine_next = j;
    }
    mine_next = warp_min_int(mine_next);
    if ((threadIdx.x & 31) == 0 && mine_next != kNone) atomicMin(&ctl->e_nx[par ^ 1], mine_next);
  }
  // last CTA to finish commits (every CTA has read npivots / e_nx[par] by then)
  __shared__ bool is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int t = atomicAdd(&ctl->ticket2, 1u);
    is_last = (t == gridDim.x - 1);
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
  plog[np % log_cap] = make_int2(e, l);
  int t = pos2var[e];
  pos2var[e] = pos2var[n + l];
  pos2var[n + l] = t;
  ctl->base.npivots = np + 1;
}

// K3 (sharded): identical streaming update on the local rows (objective row = local row mloc).
template <int kThreads, int kRowsPerCta, int kUnroll, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
ks_update(const CtlS* __restrict__ ctl, double* __restrict__ T, long long ld, int mloc, int n, int row0,
          int row1, const double* __restrict__ rowbuf2, double* col0, double* col1,
          double* __restrict__ bcol) {
  if (ctl->base.status != kRunning) return;
  const long long np = ctl->base.npivots;               // already counts the pivot in flight
  const int par = (int)(np & 1);
  const int e = ctl->base.e_cur, lg = ctl->base.l_cur, e2 = ctl->e_nx[par ^ 1];
  const int l = (lg >= row0 && lg < row1) ? lg - row0 : -1;
  const double p = ctl->base.p;
  const double* rowbuf = rowbuf2 + (long long)par * ld;
  const double* acol = ((np - 1) & 1) ? col1 : col0;
  double* ncol = (np & 1) ? col1 : col0;

  const long long j0 = ((long long)blockIdx.x * kThreads + threadIdx.x) * 4;
  if (j0 >= ld) return;
  const D4 r = *reinterpret_cast<const D4*>(rowbuf + j0);
  const int ke = (e >= j0 && e < j0 + 4) ? (int)(e - j0) : -1;
  const int k2 = (e2 != kNone && e2 >= j0 && e2 < j0 + 4) ? (int)(e2 - j0) : -1;
  const int kb = (n >= j0 && n < j0 + 4) ? (int)(n - j0) : -1;
  const bool special = (ke >= 0) | (k2 >= 0) | (kb >= 0);
  const int i_begin = blockIdx.y * kRowsPerCta;
  const int i_end = min(i_begin + kRowsPerCta, mloc + 1);
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
          o = r;
        }
        if (special) {
          if (k2 >= 0) ncol[ii] = (k2 == 0) ? o.x : (k2 == 1) ? o.y : (k2 == 2) ? o.z : o.w;
          if (kb >= 0) bcol[ii] = (kb == 0) ? o.x : (kb == 1) ? o.y : (kb == 2) ? o.z : o.w;
        }
      }
    }
  }
}

// synthetic LPs, rows [row0,row1) of the m x n instance + the replicated objective
// row; bit-identical to the numpy restatements in oracle/tier_f.py.
//   kGenDense       A_ij = u(i*n+j), b_i = (n/4)(1+u), c_j = +-u(mn+j) (positive for `param` per mille)
//   kGenUnbounded   the dense LP with every c_j > 0 and column `param` of A negated: no positive entry
//                   in that column, so the LP is unbounded and the loop says so when the column enters
//   kGenAssignment  degenerate: 0/1 incidence matrix of a bipartite graph (L = m/2 left rows, the rest
//                   right rows; column j joins left row j % L and right row L + (j % L + 7919 (j / L)) % R),
//                   b = 1, c = 1.  Totally unimodular: every tableau entry stays in {-1,0,1}, exact in
//                   binary64 and in the reference's decimal arithmetic, with many zero-ratio ties.
enum GenKind : int { kGenDense = 0, kGenUnbounded = 1, kGenAssignment = 2 };

__global__ void ks_generate_lp(double* T, long long ld, int m, int n, int row0, int row1,
                               unsigned long long seed, int kind, int param) {
  const unsigned long long mn = (unsigned long long)m * (unsigned long long)n;
  const int mloc = row1 - row0;
  const int L = m / 2, R = m - L;
  for (int il = blockIdx.y; il <= mloc; il += gridDim.y) {
    double* row = T + (long long)il * ld;
    const int i = row0 + il;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < ld;
         j += (long long)gridDim.x * blockDim.x) {
      double v = 0.0;
      if (kind == kGenAssignment) {
        if (il < mloc) {
          if (j < n) {
            const int left = (int)(j % L);
            const int right = L + (int)(((j % L) + 7919ll * (j / L)) % R);
            v = (i == left || i == right) ? 1.0 : 0.0;
          } else if (j == n) {
            v = 1.0;
          }
        } else if (j < n) {
          v = 1.0;
        }
      } else if (il < mloc) {
        if (j < n) {
          v = synth_u(seed, (unsigned long long)i * n + j);
          if (kind == kGenUnbounded && j == param) v = -v;
        } else if (j == n) {
          v = __dmul_rn((double)n / 4.0, __dadd_rn(1.0, synth_u(seed, mn + n + i)));
        }
      } else if (j < n) {
        v = synth_u(seed, mn + j);
        if (kind == kGenDense && param < 1000) {
          unsigned long long sel = splitmix64(seed ^ ~(unsigned long long)j) % 1000ULL;
          if (sel >= (unsigned long long)param) v = -v;
        }
      }
      row[j] = v;
    }
  }
}

}  // namespace lps
