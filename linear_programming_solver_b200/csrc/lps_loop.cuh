// Persistent pivot-loop kernel: ONE cooperative launch executes the whole loop of
// LPSolver.simplex (LPSolver.java:101-112) — getEntering / getLeaving / pivot until a verdict or
// the pivot cap — with software grid barriers between the three phases of a pivot instead of
// three dependent kernel launches.  The dependency chain of a pivot (column -> ratio -> row ->
// update) is serial, so at small tableaus and at 8-way sharding the launch gaps dominate; here a
// pivot costs three grid barriers (~1-2 us each) on top of the memory pass.
//
// Phases (every CTA of the co-resident grid takes part in all of them):
//   A  ratio test over the staged entering column (LPState.java:287-305), per-CTA partials,
//      [barrier 1], every CTA reduces the partials to the same (ratio,row) winner; sharded: CTA 0
//      pushes the rank's candidate into every peer's mailbox and all CTAs acquire the G candidates
//   B  pivot row (LPState.java:137-146): r_j = A[l][j]/p in 128-column chunks, written to T[l] and
//      the row buffer (sharded owner: also into every peer's row buffer + per-chunk release flag;
//      others acquire the flag), fused with the first-positive scan of the NEW objective row
//      (LPState.java:274-285) -> atomicMin, [barrier 2]
//   C  the tableau update (LPState.java:150-178) over 32-row x 512-column tiles, 256-bit
//      accesses, emitting the next entering column and the b column, [barrier 3]
// Cross-CTA data (staging vectors, control words, the tableau) is read with L1-bypassing
// `ld.global.cg` (LDG.STRONG.GPU) because the L1 is not coherent within one launch.
// Arithmetic is the same separately-rounded mul / sub / div as the multi-kernel path, so the
// two paths (and the CPU twin) agree bit for bit.
#pragma once
#include "lps_sharded.cuh"

namespace lps {

constexpr int kGroupThreads = 128;   // one tile-streaming group = 4 warps = 512 columns
constexpr int kLoopRows = 32;
constexpr int kLoopChunk = 128;       // phase-B columns per CTA step == sharded flag granularity
static_assert(kMaxChunks * kLoopChunk >= 524288, "flag slots");

struct LoopArgs {
  CtlS* ctl;
  double* T;
  long long ld;
  int mloc, n, row0, row1;
  double *col0, *col1, *bcol;
  double* rowbuf;                     // single GPU: [ld]
  Cand* partials;
  int2* plog;
  long long log_cap;
  int* pos2var;
  double eps, inf;
  Peers peers;
  int rank, world;
};

__device__ __forceinline__ double ldcg(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ldcg(const int* p) {
  int v;
  asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ D4 ldcg256(const double* p) {
  D4 v;
  asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// All CTAs of the (co-resident, cooperatively launched) grid.  The counter is monotonic and
// `target` lives in every thread.  Returns false if some CTA raised ctl->abort (a peer rank timed
// out): then nobody may wait for anybody any more and every CTA leaves the kernel.
__device__ __forceinline__ bool grid_barrier(CtlS* ctl, unsigned long long& target) {
  __shared__ int s_alive;
  target += gridDim.x;
  __syncthreads();
  if (threadIdx.x == 0) {
    int alive = 1;
    __threadfence();
    atomicAdd(&ctl->bar, 1ull);
    unsigned int spins = 0;
    while (ld_acquire_gpu_u64(&ctl->bar) < target) {
      if ((++spins & 1023u) == 0 && ldcg(&ctl->abort) != 0) { alive = 0; break; }
    }
    __threadfence();
    s_alive = alive;
  }
  __syncthreads();
  return s_alive != 0;
}

// One CTA per SM (grid barriers then cost one arrival per SM); the CTA is kGroups independent
// 128-thread groups for the streaming phase, each claiming its own tiles.
template <bool kSharded, int kLoopUnroll, int kGroups>
__global__ void __launch_bounds__(kGroupThreads * kGroups, 1) k_loop(const LoopArgs a) {
  constexpr int kLoopThreads = kGroupThreads * kGroups;
  __shared__ Cand sh_c[kLoopThreads / 32];
  __shared__ int sh_i[kLoopThreads / 32];
  __shared__ double s_slack, s_p;
  __shared__ int s_row, s_ok;
  __shared__ int s_ok2[kGroups];

  CtlS* const ctl = a.ctl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x, G = gridDim.x;
  const long long ld = a.ld;
  const int mloc = a.mloc, n = a.n;
  unsigned long long bar_target = 0;

  long long np = ctl->base.npivots;
  const long long limit = ctl->base.pivot_limit;
  int e = ctl->e_nx[(np + 1) & 1];
  const int nchunks = (int)((ld + kLoopChunk - 1) / kLoopChunk);
  const int tiles_x = (int)((ld + 4 * kGroupThreads - 1) / (4 * kGroupThreads));
  const int group = tid / kGroupThreads, gtid = tid % kGroupThreads;
  const int nstreams = G * kGroups;          // independent tile consumers in the grid
  // tile height: 32 rows when that still gives every CTA several tiles, else 16, else 8, so that
  // small (L2-resident) tableaus spread over the whole grid
  int tile_rows = kLoopRows;
  while (tile_rows > kLoopUnroll &&
         (long long)tiles_x * ((mloc + tile_rows) / tile_rows) < 4ll * nstreams) tile_rows >>= 1;
  const int tiles_y = (mloc + 1 + tile_rows - 1) / tile_rows;
  const int ntiles = tiles_x * tiles_y;
  __shared__ int s_tile[kGroups];

  for (;;) {
    const unsigned int seq = (unsigned int)(np + 1);
    const int par = seq & 1;
    double* const col = (np & 1) ? a.col1 : a.col0;      // this pivot's entering column (old values)
    double* const ncol = (np & 1) ? a.col0 : a.col1;     // next pivot's entering column
    if (cta == 0 && tid == 0) {
      ctl->e_nx[par ^ 1] = kNone;  // atomicMin target of phase B
      ctl->tile_ctr = nstreams;    // phase C: the first tile of every group is pre-assigned
    }

    // ---------------- phase A: ratio test ----------------
    Cand best;
    best.slack = a.inf; best.row = kNone; best.pad_ = 0;
    for (int i = cta * kLoopThreads + tid; i < mloc; i += G * kLoopThreads) {
      double ai = ldcg(col + i);
      if (!(ai < a.eps)) {
        double s = __ddiv_rn(ldcg(a.bcol + i), ai);
        if (s < best.slack) { best.slack = s; best.row = i; }
      }
    }
    best = warp_cand_min(best);
    if (lane == 0) sh_c[warp] = best;
    __syncthreads();
    if (tid == 0) {
      Cand c = sh_c[0];
#pragma unroll
      for (int w = 1; w < kLoopThreads / 32; w++) c = cand_min(c, sh_c[w]);
      a.partials[cta] = c;
    }
    if (!grid_barrier(ctl, bar_target)) return;
    if (warp == 0) {
      Cand c;
      c.slack = a.inf; c.row = kNone; c.pad_ = 0;
      for (int k = lane; k < G; k += 32) {
        Cand o;
        o.slack = ldcg(&a.partials[k].slack);
        o.row = ldcg(&a.partials[k].row);
        o.pad_ = 0;
        c = cand_min(c, o);
      }
      c = warp_cand_min(c);
      double slack = c.slack, p = 0.0;
      int row = c.row;   // local row
      int ok = 1;
      if (!kSharded) {
        if (lane == 0 && row != kNone) p = ldcg(col + row);
      } else {
        const int rowg = (row == kNone) ? kNone : a.row0 + row;
        if (cta == 0 && lane < a.world) {
          PeerCand* dst = &a.peers.blk[lane]->cand[par][a.rank];
          dst->slack = slack;
          dst->p = (row == kNone) ? 0.0 : ldcg(col + row);
          dst->row = rowg;
          __threadfence_system();
          st_release_sys(&dst->seq, seq);
        }
        PeerCand pc;
        pc.slack = a.inf; pc.row = kNone; pc.p = 0.0;
        if (lane < a.world) {
          const PeerCand* src = &a.peers.blk[a.rank]->cand[par][lane];
          ok = spin_until(&src->seq, seq) ? 1 : 0;
          pc.slack = ld_volatile_f64(&src->slack);
          pc.p = ld_volatile_f64(&src->p);
          pc.row = ld_volatile_s32(&src->row);
        }
        ok = __all_sync(0xffffffffu, ok) ? 1 : 0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          double os = __shfl_xor_sync(0xffffffffu, pc.slack, off);
          double op = __shfl_xor_sync(0xffffffffu, pc.p, off);
          int orow = __shfl_xor_sync(0xffffffffu, pc.row, off);
          if (os < pc.slack || (os == pc.slack && orow < pc.row)) { pc.slack = os; pc.p = op; pc.row = orow; }
        }
        slack = pc.slack; p = pc.p; row = pc.row;   // GLOBAL row from here on
      }
      if (lane == 0) { s_slack = slack; s_p = p; s_row = row; s_ok = ok; }
    }
    __syncthreads();
    const int l = (s_row == kNone) ? -1 : s_row;   // global row (== local on a single GPU)
    const double p = s_p;
    int verdict = kRunning;
    if (kSharded && !s_ok) verdict = kCommTimeout;
    else if (e == kNone) verdict = kOptimal;          // getEntering() == -1   LPSolver.java:101
    else if (l < 0) verdict = kUnbounded;             // getLeaving() == -1    LPSolver.java:103
    else if (np >= limit) verdict = kPivotCap;
    if (verdict != kRunning) {
      if (cta == 0 && tid == 0) {
        ctl->base.status = verdict;
        ctl->base.e_cur = (e == kNone) ? -1 : e;
        ctl->base.l_cur = (verdict == kPivotCap) ? l : -1;
        ctl->base.npivots = np;
        ctl->e_nx[par] = e;
      }
      return;
    }
    if (cta == 0 && tid == 0) {                        // commit pivot(e, l)
      a.plog[np % a.log_cap] = make_int2(e, l);
      int t = a.pos2var[e];                            // exchangeIndexes, LPState.java:311-320
      a.pos2var[e] = a.pos2var[n + l];
      a.pos2var[n + l] = t;
      ctl->base.e_cur = e;
      ctl->base.l_cur = l;
      ctl->base.p = p;
    }

    // ---------------- phase B: pivot row + next entering column ----------------
    const bool i_own = !kSharded || (l >= a.row0 && l < a.row1);
    const int l_loc = l - a.row0;
    const double ce = ldcg(col + mloc);
    bool dead = false;
    double* const rb = kSharded ? a.peers.rowbuf[a.rank] + (long long)par * ld : a.rowbuf;
    int mine = kNone;
    // every 128-thread group takes its own 128-column chunks (the sharded flag granularity)
    for (int c = cta * kGroups + group; c < nchunks; c += nstreams) {
      const long long j = (long long)c * kLoopChunk + gtid;
      double r = 0.0;
      if (i_own) {
        if (j < ld) {
          if (j <= n) {
            double* tl = a.T + (long long)l_loc * ld;
            r = (j == e) ? __ddiv_rn(1.0, p) : __ddiv_rn(ldcg(tl + j), p);
            tl[j] = r;
          }
          if (kSharded) {
            for (int k = 0; k < a.world; k++) (a.peers.rowbuf[k] + (long long)par * ld)[j] = r;
          } else {
            rb[j] = r;
          }
        }
        if (kSharded) {
          __threadfence_system();
          asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(kGroupThreads) : "memory");
          if (gtid < a.world && gtid != a.rank) st_release_sys(&a.peers.blk[gtid]->row_flag[par][c], seq);
        }
      } else {
        if (gtid == 0) s_ok2[group] = spin_until(&a.peers.blk[a.rank]->row_flag[par][c], seq) ? 1 : 0;
        asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(kGroupThreads) : "memory");
        if (!s_ok2[group]) {       // a peer died: raise abort so CTAs parked in a barrier leave too
          if (gtid == 0) {
            ctl->base.status = kCommTimeout;
            ctl->abort = 1;
            __threadfence();
          }
          dead = true;
          break;
        }
        if (j < ld) r = ld_volatile_f64(rb + j);
        asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(kGroupThreads) : "memory");
      }
      if (j < n) {
        double cj = ldcg(a.T + (long long)mloc * ld + j);
        double cn = (j == e) ? -__ddiv_rn(ce, p) : __dsub_rn(cj, __dmul_rn(ce, r));
        if (cn > a.eps && (int)j < mine) mine = (int)j;
      }
    }
    if (kSharded && __syncthreads_or(dead ? 1 : 0)) return;
    mine = warp_min_int(mine);
    if (lane == 0) sh_i[warp] = mine;
    __syncthreads();
    if (tid == 0) {
      int v = sh_i[0];
#pragma unroll
      for (int w = 1; w < kLoopThreads / 32; w++) v = min(v, sh_i[w]);
      if (v != kNone) atomicMin(&ctl->e_nx[par ^ 1], v);
    }
    if (!grid_barrier(ctl, bar_target)) return;
    const int e2 = ldcg(&ctl->e_nx[par ^ 1]);
    unsigned long long t_c0 = 0;
    if (cta == 0 && tid == 0) t_c0 = globaltimer_ns();

    // ---------------- phase C: tableau update ----------------
    const int l_skip = i_own ? l_loc : -1;
    // Dynamic tile queue: CTAs finish tiles at different rates (DRAM channel / die distance), so
    // tiles beyond the first wave are claimed with an atomic counter, one claim ahead.
    int t = cta * kGroups + group;
    while (t < ntiles) {
      if (gtid == 0) s_tile[group] = atomicAdd(&ctl->tile_ctr, 1);   // next tile, fetched while this one streams
      const int tx = t % tiles_x, ty = t / tiles_x;
      const long long j0 = ((long long)tx * kGroupThreads + gtid) * 4;
      if (j0 < ld) {
        const D4 r = ldcg256(rb + j0);
        const int ke = (e >= j0 && e < j0 + 4) ? (int)(e - j0) : -1;
        const int k2 = (e2 != kNone && e2 >= j0 && e2 < j0 + 4) ? (int)(e2 - j0) : -1;
        const int kb = (n >= j0 && n < j0 + 4) ? (int)(n - j0) : -1;
        const bool special = (ke >= 0) | (k2 >= 0) | (kb >= 0);
        const int i_begin = ty * tile_rows;
        const int i_end = min(i_begin + tile_rows, mloc + 1);
        double* base = a.T + j0;
        for (int i = i_begin; i < i_end; i += kLoopUnroll) {
          D4 tv[kLoopUnroll];
          double av[kLoopUnroll];
#pragma unroll
          for (int u = 0; u < kLoopUnroll; u++) {
            int ii = i + u;
            if (ii < i_end) {
              av[u] = ldcg(col + ii);
              tv[u] = ld256(base + (long long)ii * ld);   // T is never L1-allocated: cannot be stale
            }
          }
#pragma unroll
          for (int u = 0; u < kLoopUnroll; u++) {
            int ii = i + u;
            if (ii < i_end) {
              D4 o;
              if (ii != l_skip) {
                o.x = __dsub_rn(tv[u].x, __dmul_rn(av[u], r.x));
                o.y = __dsub_rn(tv[u].y, __dmul_rn(av[u], r.y));
                o.z = __dsub_rn(tv[u].z, __dmul_rn(av[u], r.z));
                o.w = __dsub_rn(tv[u].w, __dmul_rn(av[u], r.w));
                if (ke >= 0) {
                  double q = -__ddiv_rn(av[u], p);
                  if (ke == 0) o.x = q; else if (ke == 1) o.y = q; else if (ke == 2) o.z = q; else o.w = q;
                }
                st256(base + (long long)ii * ld, o);
              } else {
                o = r;
              }
              if (special) {
                if (k2 >= 0) ncol[ii] = (k2 == 0) ? o.x : (k2 == 1) ? o.y : (k2 == 2) ? o.z : o.w;
                if (kb >= 0) a.bcol[ii] = (kb == 0) ? o.x : (kb == 1) ? o.y : (kb == 2) ? o.z : o.w;
              }
            }
          }
        }
      }
      asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(kGroupThreads) : "memory");
      t = s_tile[group];
      asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(kGroupThreads) : "memory");
    }
    if (!grid_barrier(ctl, bar_target)) return;
    np += 1;
    e = e2;
    if (cta == 0 && tid == 0) {
      ctl->base.npivots = np;
      ctl->upd_ns += globaltimer_ns() - t_c0;
    }
  }
}

}  // namespace lps
