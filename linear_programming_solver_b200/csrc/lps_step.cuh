// Look-ahead blocked loop: the panel of block k+1 runs CONCURRENTLY with the pass of block k.
//
// The serial blocked loop (lps_blocked.cuh) alternates  panel(k) -> pass(k) -> panel(k+1) -> ... ; the
// panel is a latency chain (two grid-wide decisions and, when sharded, two NVLink hops per pivot) that
// leaves the memory system idle, and on small shards it is half of the pivot.  Here one cooperative
// launch per block does both at once, on disjoint sets of SMs:
//
//   CTAs [0, P)      panel role: decides the pivots of block k+1.  The tableau it needs is "T after
//                    block k", which nobody has written yet — so it reads T_cur (the tableau BEFORE
//                    block k, which this launch's pass only reads) and replays block k's pivots and
//                    then its own on every cell it touches.  Same operations in the same order as
//                    the pass applies them, hence the same bits (LPState.java:137-178).
//   CTAs [P, grid)   pass role (lps_sweep.cuh): T_next <- T_cur with block k applied, out of place,
//                    so the panel's reads of T_cur never race with the pass's writes.
//
// The buffers swap roles every launch (ctl->cur_at[launch parity] names the current one); the pending
// pivots live in two sets that alternate the same way (the pass applies set q while the panel fills
// set q^1).  Between two launches there is no host decision: a launch whose pass has nothing pending,
// or whose panel finds the run finished, simply idles that role.
//
// Compared with kb_panel the panel role also stops re-deriving the b column and the objective row
// from the tableau for every pivot: both are carried along as running vectors (bvec, cvec), updated
// by one multiply-subtract per entry per pivot — the very operation the reference applies
// (LPState.java:164, :177) — so a pivot replays only ONE column (the entering one) and, on the owner
// of the leaving row, ONE row.
#pragma once
#include "lps_sweep.cuh"

namespace lps {

constexpr int kLookMax = 2 * kPanelMax;       // previous block + own block

struct StepArgs {
  SweepArgs sw;                    // the pass (sw.ctl, sw.Tbuf, sw.ld, sw.q are shared with the panel)
  int mloc, n, row0, row1;
  int panel_ctas;                  // CTAs [0, panel_ctas) run the panel
  int block;                       // pivots per block
  double* Acols;                   // [2][block][apitch]: pending columns of set 0 / 1
  long long apitch;
  double* bvec;                    // [mloc + 1] running b column of the local rows (+ the objective slot: -v)
  double* cvec;                    // [ld]       running objective row
  double eps, inf;
  PeerCand* partials;              // [panel_ctas * kSlotStride]
  unsigned long long* mins;        // [panel_ctas * kMinStride]
  unsigned int* syncw;             // ticket A, go A, ticket B, go B, go W: one 128-byte line each
  PeerCand* gwin;
  long long ll_off;                // sharded: packet area inside every rank's exchange block
  Peers peers;                     // rowbuf[k]: [2][block][ld] pending rows of set 0 / 1 on rank k
  int rank, world;
  int2* plog;
  long long log_cap;
  int* pos2var;
  unsigned int tag0;
  int hints;                       // L2 eviction hints (env LPS_L2_HINTS, default off: no measurable effect, profiles/r02_summary.md):
                                   // bit 0 / 1: the pass's stream loads / stores evict_first
  int cells;                       // rows / columns per panel thread and trip: 2 (16-byte staging slots) or 1 (8-byte slots)
  int stage_doubles;               // doubles of dynamic shared memory the panel role may use as operand staging
};

// arrive-and-wait of the `ncta` panel CTAs (see panel_sync in lps_blocked.cuh).  The waiter issues one
// gpu-scope acquire fence after it has seen the go word: everything the other CTAs wrote before
// their (fenced) ticket is then visible to every thread of this CTA after the closing bar.sync.
template <class Grp>
__device__ __forceinline__ bool step_sync(CtlS* ctl, unsigned int* counter, unsigned int* go, unsigned int tag,
                                          unsigned int ncta) {
  __shared__ int s_alive_step;
  Grp::sync();
  if (Grp::tid() == 0) {
    int alive = 1;
    __threadfence();
    const unsigned int old = atomicAdd(counter, 1u);
    if (old == ncta - 1) {
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
    fence_acq_rel_gpu();
    s_alive_step = alive;
  }
  Grp::sync();
  return s_alive_step != 0;
}

// lexicographic (ratio, row) minimum across a warp, carrying the pivot element
__device__ __forceinline__ void warp_peer_min(PeerCand& c) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double os = __shfl_xor_sync(0xffffffffu, c.slack, off);
    const double op = __shfl_xor_sync(0xffffffffu, c.p, off);
    const int orow = __shfl_xor_sync(0xffffffffu, c.row, off);
    if (os < c.slack || (os == c.slack && orow < c.row)) { c.slack = os; c.p = op; c.row = orow; }
  }
}

// The panel of one block, run by CTAs [0, a.panel_ctas) with NT threads each.  s_op: dynamic shared
// memory, kLookMax * NT doubles.
template <bool kSharded, int NT, class Grp = WholeCta>
__device__ __forceinline__ void panel_role(const StepArgs& a, double* s_op) {
  constexpr int NW = NT / 32;
  CtlS* const ctl = a.sw.ctl;
  if (ctl->base.status != kRunning) return;
  __shared__ double s_p[kLookMax], s_re[kLookMax], s_al[kLookMax];
  __shared__ int s_l[kLookMax], s_e[kLookMax];
  __shared__ const double* s_ap[kLookMax];      // pending column u (length mloc + 1)
  __shared__ const double* s_rp[kLookMax];      // pending row u    (length ld)
  __shared__ PeerCand s_red[NW], s_red2[NW];
  __shared__ int s_min[NW], s_min2[NW];
  __shared__ double s_slack, s_pw, s_ce, s_rn;
  __shared__ int s_row, s_ok, s_e2;
  const int tid = Grp::tid(), lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x, G = a.panel_ctas;
  const long long ld = a.sw.ld;
  const int mloc = a.mloc, n = a.n;
  const int q = a.sw.q, w = q ^ 1;
  const double* const T = a.sw.Tbuf[ctl->cur_at[q]];
  long long np = ctl->base.npivots;
  const long long limit = ctl->base.pivot_limit;
  const int tp = ctl->blk_fill[q];              // the block this launch's pass applies: complete, replayed first
                                                // (blk_fill, not blk_pend: the pass clears blk_pend[q] when it is done)
  int t = 0;                                    // my own block (set w) starts empty
  int e = ctl->e_nx[(np + 1) & 1];
  double* const rows_q = a.peers.rowbuf[a.rank] + (size_t)q * a.block * ld;
  double* const rows_w = a.peers.rowbuf[a.rank] + (size_t)w * a.block * ld;
  double* const acols_q = a.Acols + (size_t)q * a.block * a.apitch;
  double* const acols_w = a.Acols + (size_t)w * a.block * a.apitch;
  if (tid < tp) {
    s_l[tid] = ctl->blk_l2[q][tid];
    s_e[tid] = ctl->blk_e2[q][tid];
    s_p[tid] = ctl->blk_p2[q][tid];
    s_ap[tid] = acols_q + (size_t)tid * a.apitch;
    s_rp[tid] = rows_q + (size_t)tid * ld;
  }
  // my share of the rows (objective row included) and of the columns: contiguous ranges
  const int RW = ((mloc + 1 + G - 1) / G + 1) & ~1;      // even: a thread takes rows (i, i+1) with 16-byte copies
  const int ilo = cta * RW, ihi = min(ilo + RW, mloc + 1);
  const int W = (int)((((ld + G - 1) / G) + 3) / 4 * 4);
  const long long jlo = (long long)cta * W, jhi = (jlo + W < ld) ? jlo + W : ld;
  const int scribe = G - 1;
  unsigned int tag = a.tag0;
  const unsigned long long t_start = (cta == scribe && tid == 0) ? globaltimer_ns() : 0ull;
  __shared__ unsigned long long s_dbg[8];
  if (tid < 8) s_dbg[tid] = 0;
  // the panel's own clock (host: split tuning; LPS_DEBUG=1 prints it): dbg_ns[14] += duration, dbg_ns[15] += pivots
  auto clock_out = [&](int pivots_done) {
    if (cta == scribe && tid == 0 && a.sw.ncta == 0) {
      // a launch without pass CTAs (the first of a run: nothing is pending yet, so every CTA works on the panel):
      // the bookkeeping the last pass CTA would do
      ctl->cur_at[q ^ 1] = ctl->cur_at[q];
    }
    if (cta == scribe && tid == 0) {
      const int slot = (a.sw.ncta == 0) ? 12 : 14;      // [12], [13]: the run's first launch (every CTA on the panel)
      ctl->dbg_ns[slot] += globaltimer_ns() - t_start;
      ctl->dbg_ns[slot + 1] += (unsigned long long)pivots_done;
      if (a.sw.ncta != 0)
        for (int k = 0; k < 6; k++) ctl->dbg_ns[k] += s_dbg[k];
    }
  };
  unsigned long long t_mark = t_start;
  // phase clock of the scribe CTA's thread 0: dbg_ns[k] += time since the previous mark (printed with LPS_DEBUG=1)
  // (accumulated in shared memory: a global read-modify-write per mark sat on the scribe's critical path)
  auto mark = [&](int k) {
    if (cta == scribe && tid == 0) {
      const unsigned long long now = globaltimer_ns();
      s_dbg[k] += now - t_mark;
      t_mark = now;
    }
  };
  bool b_lag = false;        // bvec does not include the most recent pivot yet (its r[n] was not known in time)
  int l_last = -1;           // that pivot's leaving row (local, -1: not mine)
  Grp::sync();

  // bring bvec up to date with the most recent pivot (slot tp + t - 1) for my rows
  auto settle_b = [&]() {
    if (!b_lag) return;
    const double rn = s_rn;
    const double* alast = acols_w + (size_t)(t - 1) * a.apitch;      // my own entries: written by this thread
    for (int i = ilo + tid; i < ihi; i += NT) {
      const double bi = a.bvec[i];
      a.bvec[i] = (i == l_last) ? rn : __dsub_rn(bi, __dmul_rn(alast[i], rn));   // LPState.java:146 / :164
    }
  };

  while (t < a.block) {
    const unsigned int seq = (unsigned int)(np + 1);
    const int par = seq & 1;
    const int tt = tp + t;                       // pending pivots to replay
    // ---------------- phase A: entering column of the current state, ratio test ----------------
    PeerCand best;
    best.slack = a.inf; best.row = kNone; best.p = 0.0;
    if (e != kNone) {
      if (tid < tt) s_re[tid] = ldcg_f64(s_rp[tid] + e);   // r_u[e]: written by other CTAs, behind a sync
      double* const acol = acols_w + (size_t)t * a.apitch;
      const double rn = s_rn;
      // Two rows per thread and trip, their pending-column entries staged by 16-byte cp.async; a trip takes as many
      // rows as the staging area holds for tt pending pivots (at most 2 NT).
      // (a.cells == 1: one row per thread — shorter trips, so that on a sharded solve the owner's first packets
      // leave earlier and the peers' work overlaps the owner's later trips)
      {
      const int cpt = a.cells;
      const int nte = min(NT, (a.stage_doubles / cpt) / max(tt, 1));    // threads per trip: one slot of cpt doubles each per pending pivot
      const int cells = cpt * nte;
      double* const slot0 = s_op + tid * cpt;
      const int sstride = nte * cpt;
      auto slot = [&](int u) { return slot0 + u * sstride; };
      auto slot2 = [&](int u) { const double* q = slot(u); return cpt == 2 ? *reinterpret_cast<const double2*>(q) : make_double2(*q, 0.0); };
      for (int i0 = ilo; i0 < ihi; i0 += cells) {
        const int i = i0 + cpt * tid;
        const bool on = tid < nte && i < ihi;
        const bool two = on && cpt == 2 && (i + 1 < ihi);
        double xe0 = 0.0, xe1 = 0.0, b0 = 0.0, b1 = 0.0;
        if (on) {
          if (two) {
            for (int u = 0; u < tt; u++) cp_async16(slot(u), s_ap[u] + i);
            xe1 = T[(long long)(i + 1) * ld + e];
            b1 = a.bvec[i + 1];
          } else {
            for (int u = 0; u < tt; u++) cp_async8(slot(u), s_ap[u] + i);   // (an 8-byte cp.async with an L2 cache hint raised 'illegal instruction' on B200)
          }
          xe0 = T[(long long)i * ld + e];
          b0 = a.bvec[i];
        }
        cp_async_commit();
        cp_async_wait_all();
        Grp::sync();   // s_re (first trip)
        if (on) {
          if (b_lag) {     // the previous pivot's update of b, LPState.java:146 / :164
            const double2 al = slot2(tt - 1);
            b0 = (i == l_last) ? rn : __dsub_rn(b0, __dmul_rn(al.x, rn));
            a.bvec[i] = b0;
            if (two) {
              b1 = (i + 1 == l_last) ? rn : __dsub_rn(b1, __dmul_rn(al.y, rn));
              a.bvec[i + 1] = b1;
            }
          }
          for (int u = 0; u < tt; u++) {
            const double2 au = slot2(u);
            const double re = s_re[u];
            const bool ecol = (e == s_e[u]);
            if (i == s_l[u]) xe0 = re;                                          // :137-146
            else xe0 = ecol ? -ddiv_call(au.x, s_p[u]) : __dsub_rn(xe0, __dmul_rn(au.x, re));   // :157 / :162
            if (i + 1 == s_l[u]) xe1 = re;
            else xe1 = ecol ? -ddiv_call(au.y, s_p[u]) : __dsub_rn(xe1, __dmul_rn(au.y, re));
          }
          acol[i] = xe0;
          if (i < mloc && !(xe0 < a.eps)) {                                     // :294-299
            const double sl = ddiv_call(b0, xe0);
            if (sl < best.slack) { best.slack = sl; best.row = i; best.p = xe0; }
          }
          if (two) {
            acol[i + 1] = xe1;
            if (i + 1 < mloc && !(xe1 < a.eps)) {
              const double sl = ddiv_call(b1, xe1);
              if (sl < best.slack) { best.slack = sl; best.row = i + 1; best.p = xe1; }   // strict: the lower row keeps ties
            }
          }
        }
        Grp::sync();   // s_op is reused by the next trip / phase B
      }
      }
      b_lag = false;
    }
    mark(0);                                            // phase A trips
    warp_peer_min(best);
    if (lane == 0) s_red[warp] = best;
    Grp::sync();
    if (tid == 0) {
      PeerCand c = s_red[0];
      for (int k = 1; k < NW; k++) {
        const PeerCand o = s_red[k];
        if (o.slack < c.slack || (o.slack == c.slack && o.row < c.row)) c = o;
      }
      PeerCand* mine = &a.partials[cta * kSlotStride];
      mine->slack = c.slack;
      mine->p = c.p;
      mine->row = c.row;
    }
    if (!step_sync<Grp>(ctl, a.syncw + 0, a.syncw + 32, tag, G)) {
      if (tid == 0) { ctl->base.status = kCommTimeout; ctl->abort = 1; __threadfence(); }
      return;
    }
    mark(1);                                            // sync A
    {
      PeerCand c;
      c.slack = a.inf; c.row = kNone; c.p = 0.0;
      for (int k = tid; k < G; k += NT) {
        const PeerCand* src = &a.partials[k * kSlotStride];
        PeerCand o;
        o.slack = ldcg_f64(&src->slack);
        o.p = ldcg_f64(&src->p);
        o.row = ldcg_s32(&src->row);
        if (o.slack < c.slack || (o.slack == c.slack && o.row < c.row)) c = o;
      }
      warp_peer_min(c);
      if (lane == 0) s_red2[warp] = c;
      Grp::sync();
    }
    if (warp == 0) {
      PeerCand c;
      c.slack = a.inf; c.row = kNone; c.p = 0.0;
      if (lane < NW) c = s_red2[lane];
      warp_peer_min(c);
      int ok = 1;
      if (kSharded) {
        // only CTA 0 talks to the peers; it hands the cross-rank winner to the other CTAs through one go word
        if (cta == 0) {
          if (lane < a.world) {
            LLPacket* dst = reinterpret_cast<LLPacket*>(reinterpret_cast<char*>(a.peers.blk[lane]) + a.ll_off) +
                            2 * ld + ((size_t)par * kMaxRanks + a.rank) * 4;
            ll_store(dst + 0, c.slack, seq);
            ll_store(dst + 1, c.p, seq);
            ll_store(dst + 2, (double)((c.row == kNone) ? -1 : a.row0 + c.row), seq);
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
          warp_peer_min(pc);
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
            fence_acq_rel_gpu();
            c.slack = ldcg_f64(&a.gwin->slack);
            c.p = ldcg_f64(&a.gwin->p);
            c.row = ldcg_s32(&a.gwin->row);
          }
        }
      }
      if (lane == 0) { s_slack = c.slack; s_pw = c.p; s_row = c.row; s_ok = ok; }
    }
    Grp::sync();
    const int l = (s_row == kNone) ? -1 : s_row;       // global row (== local on a single GPU)
    const double p = s_pw;
    mark(2);                                            // gather of the partials + cross-rank candidate exchange
    if (kSharded && !s_ok) {
      if (tid == 0) { ctl->base.status = kCommTimeout; ctl->abort = 1; __threadfence(); }
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
      clock_out(t);
      return;                                           // (bvec is rebuilt from the tableau when the next run starts)
    }
    // ---------------- phase B: leaving row, objective row, next entering column ----------------
    const bool i_own = !kSharded || (l >= a.row0 && l < a.row1);
    const int lloc = i_own ? l - a.row0 : -1;
    if (tid < tt) s_al[tid] = i_own ? ldcg_f64(s_ap[tid] + lloc) : 0.0;
    if (tid == NT - 1) s_ce = ldcg_f64(acols_w + (size_t)t * a.apitch + mloc);
    int mine_next = kNone;
    {
    const int cpt = a.cells;
    const int nte = min(NT, (a.stage_doubles / cpt) / max(tt, 1));
    const int cells = cpt * nte;
    double* const slot0 = s_op + tid * cpt;
      const int sstride = nte * cpt;
      auto slot = [&](int u) { return slot0 + u * sstride; };
    auto slot2 = [&](int u) { const double* q = slot(u); return cpt == 2 ? *reinterpret_cast<const double2*>(q) : make_double2(*q, 0.0); };
    LLPacket* const ll_mine = reinterpret_cast<LLPacket*>(reinterpret_cast<char*>(a.peers.blk[a.rank]) + a.ll_off) + (size_t)par * ld;
    for (long long jb = jlo; jb < jhi; jb += cells) {      // columns j (and j+1: jlo, jhi and ld are even)
      const long long j = jb + cpt * tid;
      const bool on = tid < nte && j < jhi;
      const bool two = on && cpt == 2;
      double2 cj = make_double2(0.0, 0.0), x = make_double2(0.0, 0.0);
      if (on) {
        if (i_own) {
          if (two) {
            for (int u = 0; u < tt; u++) cp_async16(slot(u), s_rp[u] + j);
            x = *reinterpret_cast<const double2*>(T + (long long)lloc * ld + j);    // (columns past n hold zeros)
          } else {
            for (int u = 0; u < tt; u++) cp_async8(slot(u), s_rp[u] + j);
            x.x = T[(long long)lloc * ld + j];
          }
        }
        if (two) cj = *reinterpret_cast<const double2*>(a.cvec + j);
        else cj.x = a.cvec[j];
      }
      cp_async_commit();
      cp_async_wait_all();
      Grp::sync();     // s_al / s_ce
      bool got = true;
      if (on) {
        const double ce = s_ce;
        double2 r = make_double2(0.0, 0.0);
        if (i_own) {
          for (int u = 0; u < tt; u++) {
            const double2 ru = slot2(u);
            if (lloc == s_l[u]) x = ru;                       // the row was pending pivot u's leaving row
            else {
              const double al = s_al[u];
              x.x = ((int)j == s_e[u]) ? -ddiv_call(al, s_p[u]) : __dsub_rn(x.x, __dmul_rn(al, ru.x));
              if (two) x.y = ((int)j + 1 == s_e[u]) ? -ddiv_call(al, s_p[u]) : __dsub_rn(x.y, __dmul_rn(al, ru.y));
            }
          }
          if (j <= n) r.x = ((int)j == e) ? ddiv_call(1.0, p) : ddiv_call(x.x, p);      // LPState.java:139-146
          if (two && j + 1 <= n) r.y = ((int)j + 1 == e) ? ddiv_call(1.0, p) : ddiv_call(x.y, p);
          if (kSharded) {      // compute + broadcast in one kernel: packets straight into every peer's memory
            for (int k = 0; k < a.world; k++)
              if (k != a.rank) {
                LLPacket* dst = reinterpret_cast<LLPacket*>(reinterpret_cast<char*>(a.peers.blk[k]) + a.ll_off) +
                                (size_t)par * ld + j;
                ll_store(dst, r.x, seq);
                if (two) ll_store(dst + 1, r.y, seq);
              }
          }
        } else {
          got = ll_load(ll_mine + j, seq, r.x);
          if (two) got = ll_load(ll_mine + j + 1, seq, r.y) && got;
        }
        double2 cn = cj;
        if (j < n) {
          cn.x = ((int)j == e) ? -ddiv_call(ce, p) : __dsub_rn(cj.x, __dmul_rn(ce, r.x));   // :170-178
          if (cn.x > a.eps && (int)j < mine_next) mine_next = (int)j;
        }
        if (two && j + 1 < n) {
          cn.y = ((int)j + 1 == e) ? -ddiv_call(ce, p) : __dsub_rn(cj.y, __dmul_rn(ce, r.y));
          if (cn.y > a.eps && (int)j + 1 < mine_next) mine_next = (int)j + 1;
        }
        if (two) {
          *reinterpret_cast<double2*>(rows_w + (size_t)t * ld + j) = r;      // this rank's copy of the pending row
          *reinterpret_cast<double2*>(a.cvec + j) = cn;
        } else {
          rows_w[(size_t)t * ld + j] = r.x;
          a.cvec[j] = cn.x;
        }
      }
      if (Grp::sync_or(got ? 0 : 1)) {   // (also: s_op is reused by the next trip / phase A)
        if (tid == 0) { ctl->base.status = kCommTimeout; ctl->abort = 1; __threadfence(); }
        return;
      }
    }
    }
    mark(3);                                            // phase B trips (row replay / packets, objective row)
    mine_next = warp_min_int(mine_next);
    if (lane == 0) s_min[warp] = mine_next;
    if (tid == 0) {          // every CTA keeps its own copy of the pending pivots' scalars
      s_e[tt] = e;
      s_l[tt] = lloc;
      s_p[tt] = p;
      s_ap[tt] = acols_w + (size_t)t * a.apitch;
      s_rp[tt] = rows_w + (size_t)t * ld;
    }
    Grp::sync();
    if (tid == 0) {
      int v = s_min[0];
      for (int k = 1; k < NW; k++) v = min(v, s_min[k]);
      a.mins[cta * kMinStride] = (unsigned long long)(unsigned int)v;
    }
    if (!step_sync<Grp>(ctl, a.syncw + 64, a.syncw + 96, tag, G)) {
      if (tid == 0) { ctl->base.status = kCommTimeout; ctl->abort = 1; __threadfence(); }
      return;
    }
    mark(4);                                            // sync B
    {
      int v = kNone;
      for (int k = tid; k < G; k += NT) {
        unsigned long long wv;
        asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(wv) : "l"(a.mins + k * kMinStride) : "memory");
        v = min(v, (int)(unsigned int)(wv & 0xffffffffull));
      }
      v = warp_min_int(v);
      if (lane == 0) s_min2[warp] = v;
      if (tid == NT - 1) s_rn = ldcg_f64(rows_w + (size_t)t * ld + n);     // b_l / p of this pivot: the next b update
      Grp::sync();
      if (tid == 0) {
        int m2 = s_min2[0];
        for (int k = 1; k < NW; k++) m2 = min(m2, s_min2[k]);
        s_e2 = m2;
      }
      Grp::sync();
    }
    const int e2 = s_e2;
    if (cta == scribe && tid == 0) {                       // commit pivot(e, l)
      ctl->base.e_cur = e;
      ctl->base.l_cur = l;
      ctl->base.p = p;
      ctl->owner = i_own ? a.rank : -1;
      ctl->blk_e2[w][t] = e;
      ctl->blk_l2[w][t] = lloc;
      ctl->blk_p2[w][t] = p;
      ctl->blk_pend[w] = t + 1;
      ctl->blk_fill[w] = t + 1;
      ctl->e_nx[par ^ 1] = e2;
      a.plog[np % a.log_cap] = make_int2(e, l);
      const int tmp = a.pos2var[e];                        // exchangeIndexes, LPState.java:311-320
      a.pos2var[e] = a.pos2var[n + l];
      a.pos2var[n + l] = tmp;
      ctl->base.npivots = np + 1;
    }
    np += 1;
    t += 1;
    e = e2;
    tag += 1;
    b_lag = true;
    l_last = lloc;
    mark(5);                                            // gather of the minima + commit
  }
  settle_b();
  clock_out(t);
}

// one block step: panel of the next block on CTAs [0, P), pass of the current block on the rest
template <bool kSharded, class Shape>
__global__ void __launch_bounds__(Shape::kThreads, 1)
kb_step(const __grid_constant__ StepArgs a, const __grid_constant__ CUtensorMap tmT0,
        const __grid_constant__ CUtensorMap tmT1, const __grid_constant__ CUtensorMap tmA,
        const __grid_constant__ CUtensorMap tmR) {
  extern __shared__ __align__(1024) unsigned char step_smem[];
  if ((int)blockIdx.x < a.panel_ctas) panel_role<kSharded, Shape::kThreads>(a, reinterpret_cast<double*>(step_smem));
  else sweep_role<Shape>(a.sw, &tmT0, &tmT1, &tmA, &tmR, step_smem);
}

// the same step with the cp.async pass (flush_role, lps_blocked.cuh) instead of the TMA pipeline
template <bool kSharded, int kLanes, int kU, int kG, bool kPre>
__global__ void __launch_bounds__(kFlushThreads * kLanes, 1)
kb_step_flush(const __grid_constant__ StepArgs a) {
  extern __shared__ __align__(1024) unsigned char step_smem[];
  if ((int)blockIdx.x < a.panel_ctas) {
    panel_role<kSharded, kFlushThreads * kLanes>(a, reinterpret_cast<double*>(step_smem));
  } else {
    FlushArgs fa;
    fa.ctl = a.sw.ctl;
    fa.Tbuf[0] = a.sw.Tbuf[0];
    fa.Tbuf[1] = a.sw.Tbuf[1];
    fa.ld = a.sw.ld;
    fa.mloc = a.mloc;
    fa.Acols = a.Acols;
    fa.apitch = a.apitch;
    fa.Rrows = a.peers.rowbuf[a.rank];
    fa.block = a.block;
    fa.q = a.sw.q;
    fa.inplace = a.sw.inplace;
    fa.ncta = a.sw.ncta;
    fa.hints = a.hints;
    flush_role<kLanes, kU, kG, kPre>(fa, reinterpret_cast<double*>(step_smem));
  }
}

// The look-ahead step with BOTH roles in every CTA: warps 0-11 run the cp.async pass (three row-group lanes),
// warps 12-15 the panel.  The panel is a latency chain — its threads mostly wait for L2 / DRAM / a grid-wide
// decision / an NVLink packet — so instead of lending it whole SMs that then sit idle (kb_step_flush), every SM
// lends it four warps whose waiting costs nothing, and the panel runs at full-GPU width: one or two trips per
// step instead of a dozen.  The two groups have their own named barriers and never wait for each other.
template <bool kSharded, int kG>
__global__ void __launch_bounds__(512, 1)
kb_step_ws(const __grid_constant__ StepArgs a) {
  extern __shared__ __align__(1024) unsigned char step_smem[];
  constexpr int kPassThreads = 384, kPanelThreads2 = 128;
  constexpr int kCH = 4 * 3 * kG;
  const size_t pass_bytes = (size_t)a.block * 2 * (kStripCols + kCH) * sizeof(double);
  if ((int)threadIdx.x < kPassThreads) {
    FlushArgs fa;
    fa.ctl = a.sw.ctl;
    fa.Tbuf[0] = a.sw.Tbuf[0];
    fa.Tbuf[1] = a.sw.Tbuf[1];
    fa.ld = a.sw.ld;
    fa.mloc = a.mloc;
    fa.Acols = a.Acols;
    fa.apitch = a.apitch;
    fa.Rrows = a.peers.rowbuf[a.rank];
    fa.block = a.block;
    fa.q = a.sw.q;
    fa.inplace = a.sw.inplace;
    fa.ncta = (int)gridDim.x;
    fa.hints = a.hints;
    flush_role<3, 4, kG, true, CtaGroup<1, kPassThreads, 0>>(fa, reinterpret_cast<double*>(step_smem));
  } else {
    panel_role<kSharded, kPanelThreads2, CtaGroup<2, kPanelThreads2, kPassThreads>>(
        a, reinterpret_cast<double*>(step_smem + pass_bytes));
  }
}

// run set-up of the look-ahead loop: the running b column and objective row start as the tableau's own
__global__ void kb_init_vec(const double* __restrict__ T, long long ld, int mloc, int n,
                            double* __restrict__ bvec, double* __restrict__ cvec) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k <= mloc) bvec[k] = T[k * ld + n];
  if (k < ld) cvec[k] = T[(long long)mloc * ld + k];
}

}  // namespace lps
