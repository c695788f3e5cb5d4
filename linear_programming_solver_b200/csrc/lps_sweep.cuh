// The tableau pass of the blocked loop as a TMA + mbarrier pipeline (sm_100a).
//
// What it computes is kb_flush's job (lps_blocked.cuh): every cell of the local tableau takes the
// pending pivots u = 0..t-1 in order,
//
//     i == l_u  ->  x = r_u[j]                 (LPState.java:137-146: row l becomes the scaled row)
//     j == e_u  ->  x = -(a_u[i] / p_u)        (:157 / :172)
//     else      ->  x = x - a_u[i] * r_u[j]    (:162-164 / :177, multiply and subtract rounded separately)
//
// How it moves the data is new:
//
//   * one PRODUCER warp (one elected lane) streams the tableau through a ring of kStages shared-memory
//     stages with cp.async.bulk.tensor (2-D FP64 tensor maps of T, of the pending columns and of the
//     pending rows); every stage is a tile of kSR rows x 256 columns plus the matching slice
//     a_u[i0 .. i0+kSR) of the pending columns.  Stages are guarded by full / empty mbarriers, so the
//     loads run kStages-1 tiles (> 100 KB per SM) ahead of the arithmetic and no warp ever waits for
//     DRAM with its registers tied up;
//   * twelve CONSUMER warps (three per scheduler) replay the pending pivots.  A thread owns TWO
//     columns of the strip for the whole chunk and keeps r_u[j], r_u[j+1] of all pending pivots in
//     REGISTERS (loaded once per chunk from a TMA-staged slice), so the inner loop reads only the
//     a_u[i] operands from shared memory — warp-wide broadcasts, one wavefront each — instead of
//     four 128-bit shared loads per pivot per row group; the x values come out of the tile with one
//     128-bit shared load per row and leave with one 128-bit global store;
//   * warps never meet at a CTA-wide barrier: a warp waits for the stage's full barrier, computes
//     its 4 rows x 64 columns, arrives on the empty barrier and moves on, so the warps of one SM
//     drift apart by up to kStages tiles and the FP64 pipe always has somebody to issue for;
//   * the pass may be out of place (src != dst): the look-ahead loop (lps_step.cuh) lets the panel of
//     the NEXT block read the old tableau while this pass writes the new one.
//
// Cells that a pending pivot overwrites (its leaving row, its entering column) are handled in
// sequence inside the replay, on a path only the warps that hold such a cell in this stage take.
#pragma once
#include <cuda.h>

#include "lps_blocked.cuh"

namespace lps {

constexpr int kSwCols = 256;                              // strip width (TMA box limit: 256 elements)
constexpr unsigned long long kSwWaitNs = 4ull * 1000ull * 1000ull * 1000ull;

// Shape of the pass: kS = most pending pivots (their rows live in registers), kR x kC = rows x columns
// per consumer thread per stage (kC = 2 or 4), kCW = consumer warps: 256 / (32 kC) of them span the
// columns of a strip, the rest are stacked over the rows of a stage.
//   Every FP64 instruction needs one a_u[i] operand, and that operand reaches every lane through the
//   shared-memory data pipe (8 bytes x 32 lanes = two wavefronts per value, broadcast or not); it is
//   reused by the thread's kC columns, so the pipe carries 8 / (2 kC) bytes per FP64 instruction per
//   lane: with kC = 2 the LSU is as busy as the FP64 pipe (measured: both stuck near 55-60 %), with
//   kC = 4 half as busy — paid for with 2 kS kC registers for the pending rows.
// The register file is handed out to CTAs in units of four warps, so 8 + 1 warps cost 12 warps' worth
// of registers (168 per thread) and 12 + 1 cost 16 (128 per thread); the padding warps idle in the
// pass and work in the panel role.
template <int kS_, int kR_, int kC_, int kCW_>
struct SweepShape {
  static constexpr int kS = kS_, kR = kR_, kC = kC_, kCW = kCW_;
  static constexpr int kColWarps = kSwCols / (32 * kC);      // warps side by side over the strip
  static constexpr int kRowLanes = kCW / kColWarps;
  static constexpr int kSR = kR * kRowLanes;                 // rows per stage
  static constexpr int kThreads = (kCW + 1 + 3) / 4 * 4 * 32;   // whole register-allocation units of four warps
  static constexpr size_t kTile = (size_t)kSR * kSwCols * sizeof(double);
  static constexpr size_t kASlice = (size_t)kS * kSR * sizeof(double);
  static constexpr size_t kRSlice = (size_t)kS * kSwCols * sizeof(double);
  static constexpr size_t kBudget = 220 * 1024;
  static constexpr int kStagesFit = (int)((kBudget - 2 * kRSlice) / (kTile + kASlice));
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  // shared-memory carve-up (dynamic shared memory, 1024-byte aligned base)
  static constexpr size_t kOffTiles = 0;
  static constexpr size_t kOffA = kOffTiles + kStages * kTile;
  static constexpr size_t kOffR = kOffA + kStages * kASlice;
  static constexpr size_t kOffBars = kOffR + 2 * kRSlice;          // the pending-row slice is double-buffered by chunk
  static constexpr size_t kOffMeta = kOffBars + (2 * kStages + 4) * sizeof(unsigned long long);
  static constexpr size_t kOffScal = kOffMeta + 2 * 16;
  static constexpr size_t kBytes = kOffScal + (size_t)kS * (8 + 4 + 4) + 64;
  static_assert((kC == 2 || kC == 4) && kCW % kColWarps == 0 && kR % 2 == 0, "shape");
  static_assert(kStages >= 3, "pipeline depth");
  static_assert(kASlice % 128 == 0 && kTile % 128 == 0, "TMA destinations are 128-byte aligned");
};

// ---- mbarrier / TMA primitives (PTX ISA 8.x, sm_90+) ------------------------------------------
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes (or the hint
// runs out) instead of returning at once — a waiting warp then costs no issue slots (without the hint the
// producer's poll loop took 15 % of its scheduler's slots away from the consumer warps next to it)
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned int parity) {
  unsigned int ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(200000u)
      : "memory");
  return ok != 0;
}
// a wait that cannot hang the GPU: a pipeline bug traps (the launch fails loudly) instead of spinning forever
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0 = 0;
  unsigned int spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 63u) == 0) {
      const unsigned long long now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kSwWaitNs) __trap();
    }
  }
}
// 2-D tile: coordinates are {innermost (column), outer (row)} in elements; out-of-range parts arrive as zeros
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                                 unsigned long long* bar, unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ double2 lds128(unsigned int saddr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void st128_stream(double* p, double x, double y) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(x), "d"(y) : "memory");
}

struct SweepArgs {
  CtlS* ctl;
  double* Tbuf[2];             // the tableau buffers, (mloc+1) x ld each; ctl->cur_at[q] names the current one.
                               // In place: both entries are the same buffer.
  long long ld;
  int rows;                    // mloc + 1 (objective row included)
  int chunk_rows;              // rows per chunk, a multiple of the stage height
  int bw;                      // tile width in columns = min(256, ld): box of the T and pending-row maps
  int bu;                      // pending pivots per box = min(kS, block_pivots)
  int q;                       // launch parity: the pass applies pending set q (ctl->blk_*2[q])
  int inplace;                 // 1: write back into the current buffer; 0: write the other one and publish the flip
  int cta0, ncta;              // CTAs [cta0, cta0 + ncta) of the grid run the pass (the others: the panel)
};

// The pass, run by CTAs [a.cta0, a.cta0 + a.ncta) with Shape::kThreads threads each.  `smem` is the CTA's
// dynamic shared memory (>= Shape::kBytes, 1024-byte aligned).
// tmT0 / tmT1: tensor maps of Tbuf[0] / Tbuf[1]; tmA / tmR: pending columns / rows of set a.q.
template <class Shape>
__device__ __forceinline__ void sweep_role(const SweepArgs& a, const CUtensorMap* tmT0, const CUtensorMap* tmT1,
                                           const CUtensorMap* tmA, const CUtensorMap* tmR, unsigned char* smem) {
  using SM = Shape;
  constexpr int kS = Shape::kS, kSwR = Shape::kR, kSwSR = Shape::kSR, kSwStages = Shape::kStages;
  constexpr int kSwConsumerWarps = Shape::kCW;
  CtlS* const ctl = a.ctl;
  const int set = a.q;
  const int t = ctl->blk_pend[set];
  const int cur = ctl->cur_at[a.q];
  const CUtensorMap* const tmT = cur ? tmT1 : tmT0;
  double* const dst = a.Tbuf[a.inplace ? cur : (cur ^ 1)];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* const tiles = reinterpret_cast<double*>(smem + SM::kOffTiles);
  double* const s_a = reinterpret_cast<double*>(smem + SM::kOffA);
  double* const s_r = reinterpret_cast<double*>(smem + SM::kOffR);
  unsigned long long* const full = reinterpret_cast<unsigned long long*>(smem + SM::kOffBars);
  unsigned long long* const empty = full + kSwStages;
  unsigned long long* const r_full = empty + kSwStages;      // [2]
  unsigned long long* const r_empty = r_full + 2;             // [2]
  int4* const desc = reinterpret_cast<int4*>(smem + SM::kOffMeta);      // chunk descriptor {j0, first row, stages, stop}
  double* const s_p = reinterpret_cast<double*>(smem + SM::kOffScal);
  int* const s_l = reinterpret_cast<int*>(s_p + kS);
  int* const s_e = s_l + kS;
  __shared__ bool s_last;

  if (t > 0) {
    if (tid < t) {
      s_l[tid] = ctl->blk_l2[set][tid];
      s_e[tid] = ctl->blk_e2[set][tid];
      s_p[tid] = ctl->blk_p2[set][tid];
    }
    if (tid == 0) {
      for (int s = 0; s < kSwStages; s++) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], kSwConsumerWarps);
      }
      for (int b2 = 0; b2 < 2; b2++) {
        mbar_init(&r_full[b2], 1);
        mbar_init(&r_empty[b2], kSwConsumerWarps);
      }
      mbar_fence_init();
    }
    __syncthreads();

    const int bw = a.bw;
    const int nstrips = (int)((a.ld + bw - 1) / bw);
    const int nbands = (a.rows + a.chunk_rows - 1) / a.chunk_rows;
    const long long nchunks = (long long)nstrips * nbands;
    const unsigned int tile_bytes = (unsigned int)(kSwSR * bw * sizeof(double));
    const unsigned int a_bytes = (unsigned int)(a.bu * kSwSR * sizeof(double));
    const unsigned int r_bytes = (unsigned int)(a.bu * bw * sizeof(double));

    if (warp > kSwConsumerWarps) {
      // padding warps of the register-allocation unit: nothing to do in the pass
    } else if (warp == kSwConsumerWarps) {
      // ---------------- producer: one lane issues every bulk copy of this CTA ----------------
      // Per chunk: the chunk descriptor {j0, first row, stages, stop} and the chunk's slice of the pending
      // rows travel under the r_full / r_empty barrier pair; then one tile + a-slice per stage under the
      // stage's full / empty pair.  The consumers count the stages of a chunk themselves.
      if (lane == 0) {
        tma_prefetch_desc(tmT);
        tma_prefetch_desc(tmA);
        tma_prefetch_desc(tmR);
        const unsigned long long pol = l2_policy_evict_first();
        unsigned long long* const queue = &ctl->blk_queue;
        int stage = 0;
        unsigned int phase = 0;
        for (unsigned int nc = 0;; nc++) {
          const long long c = (long long)atomicAdd(queue, 1ull);
          const int rb = nc & 1u;                        // slice buffer / descriptor slot of this chunk
          mbar_wait(&r_empty[rb], ((nc >> 1) & 1u) ^ 1u);   // the chunk that used this buffer two chunks ago is done
          if (c >= nchunks) {
            desc[rb] = make_int4(0, 0, 0, 1);            // stop
            mbar_arrive(&r_full[rb]);
            break;
          }
          // chunks are numbered row-band-major: the CTAs of the grid sweep the tableau as one band
          const int j0 = (int)(c % nstrips) * bw;
          const int ib = (int)(c / nstrips) * a.chunk_rows;
          const int iend = min(ib + a.chunk_rows, a.rows);
          desc[rb] = make_int4(j0, ib, (iend - ib + kSwSR - 1) / kSwSR, 0);
          mbar_arrive_expect_tx(&r_full[rb], r_bytes);
          tma_load_2d(s_r + (size_t)rb * kS * kSwCols, tmR, j0, 0, &r_full[rb]);
          for (int i0 = ib; i0 < iend; i0 += kSwSR) {
            mbar_wait(&empty[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&full[stage], tile_bytes + a_bytes);
            tma_load_2d_hint(tiles + (size_t)stage * kSwSR * kSwCols, tmT, j0, i0, &full[stage], pol);
            tma_load_2d(s_a + (size_t)stage * kS * kSwSR, tmA, i0, 0, &full[stage]);
            if (++stage == kSwStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    } else {
      // ---------------- consumers ----------------
      constexpr int kC = Shape::kC, kH = kC / 2;          // kH column pairs per thread, 64 columns apart
      const int cgrp = warp % Shape::kColWarps, rlane = warp / Shape::kColWarps;
      const int wcol = cgrp * 32 * kC;                     // first column of my warp inside the strip
      const int jt = wcol + lane * 2;                      // my first pair; the second one (kC == 4) is 64 further
      // byte offsets of my cells inside a stage's tile / a-slice: fixed for the whole pass
      const unsigned int tile_off = (unsigned int)(((rlane * kSwR) * bw + jt) * sizeof(double));
      const unsigned int tiles_s = smem_u32(tiles) + tile_off;
      const unsigned int sa_s = smem_u32(s_a) + (unsigned int)(rlane * kSwR * sizeof(double));
      const unsigned int row_b = (unsigned int)(bw * sizeof(double));
      double r[kS][kC];
#pragma unroll
      for (int u = 0; u < kS; u++)
#pragma unroll
        for (int c = 0; c < kC; c++) r[u][c] = 0.0;
      unsigned int colmask = 0;                          // pending pivots whose entering column my WARP holds
      unsigned int rowmask = 0;                          // pending pivots whose leaving row lies in this chunk
      int stage = 0;
      unsigned int phase = 0;
      for (unsigned int nc = 0;; nc++) {
        // ---- a new chunk: descriptor + my slice of the pending rows ----
        const int rb = nc & 1u;
        mbar_wait(&r_full[rb], (nc >> 1) & 1u);
        const int4 ds = desc[rb];
        if (ds.w) break;
        const int j0 = ds.x, ib = ds.y;
        int nst = ds.z;
        const double* const srb = s_r + (size_t)rb * kS * kSwCols;    // stays valid until I release it below
#pragma unroll
        for (int h = 0; h < kH; h++) {
          if (jt + 64 * h < bw) {
#pragma unroll
            for (int u = 0; u < kS; u++)
              if (u < t) {
                const double2 v = *reinterpret_cast<const double2*>(srb + (size_t)u * bw + jt + 64 * h);
                r[u][2 * h] = v.x;
                r[u][2 * h + 1] = v.y;
              }
          }
        }
        colmask = rowmask = 0;
        for (int u = 0; u < t; u++) {
          const int d = s_e[u] - (j0 + wcol);
          if (d >= 0 && d < 32 * kC) colmask |= 1u << u;
          const int dr = s_l[u] - ib;
          if (dr >= 0 && dr < a.chunk_rows) rowmask |= 1u << u;
        }
        const int j = j0 + jt;
        bool act[kH];
#pragma unroll
        for (int h = 0; h < kH; h++) act[h] = (jt + 64 * h < bw) && (j + 64 * h < a.ld);
        int i = ib + rlane * kSwR;                       // my first row of the stage
        double* out = dst + (long long)i * a.ld + j;
        // ---- the stages of the chunk ----
        bool ready = mbar_try_wait(&full[stage], phase);
        for (; nst > 0; nst--) {
          if (!ready) mbar_wait(&full[stage], phase);
          unsigned int smask = colmask;                  // ... or whose leaving row is one of my rows
          if (rowmask != 0) {
            for (int u = 0; u < t; u++) {
              const int d = s_l[u] - i;
              if (d >= 0 && d < kSwR) smask |= 1u << u;
            }
          }
          const unsigned int tile = tiles_s + (unsigned int)stage * (unsigned int)(kSwSR * kSwCols * sizeof(double));
          const unsigned int sa = sa_s + (unsigned int)stage * (unsigned int)(kS * kSwSR * sizeof(double));
          double2 x[kSwR][kH];
#pragma unroll
          for (int k = 0; k < kSwR; k++)
#pragma unroll
            for (int h = 0; h < kH; h++) x[k][h] = lds128(tile + k * row_b + 64 * h * (unsigned int)sizeof(double));
          // one pending pivot on my kSwR x kC cells: a_u[i..i+kSwR) from shared memory (warp-wide broadcasts)
          auto load_a = [&](int u, double (&av)[kSwR]) {
#pragma unroll
            for (int k = 0; k < kSwR; k += 2) {
              const double2 v = lds128(sa + (unsigned int)((u * kSwSR + k) * sizeof(double)));
              av[k] = v.x;
              av[k + 1] = v.y;
            }
          };
          auto update = [&](int u, const double (&av)[kSwR]) {
#pragma unroll
            for (int k = 0; k < kSwR; k++)
#pragma unroll
              for (int h = 0; h < kH; h++) {
                x[k][h].x = __dsub_rn(x[k][h].x, __dmul_rn(av[k], r[u][2 * h]));      // LPState.java:162-164 / :177
                x[k][h].y = __dsub_rn(x[k][h].y, __dmul_rn(av[k], r[u][2 * h + 1]));
              }
          };
          if (smask == 0 && t == kS) {
            // the common case, free of branches: the operands of pivot u + 1 are fetched while pivot u is applied
            if constexpr (Shape::kThreads >= 512) {
              // 128 registers per thread: one operand buffer (the other two warps of the scheduler cover the
              // shared-memory latency); the second buffer cost spills of the loop state
#pragma unroll
              for (int u = 0; u < kS; u++) {
                double av[kSwR];
                load_a(u, av);
                update(u, av);
              }
            } else {
              double a0[kSwR], a1[kSwR];
              load_a(0, a0);
#pragma unroll
              for (int u = 0; u < kS; u += 2) {
                load_a(u + 1, a1);
                update(u, a0);
                if (u + 2 < kS) load_a(u + 2, a0);
                update(u + 1, a1);
              }
            }
          } else {
            // a partial block, or a pending pivot OVERWRITES some of my warp's cells in this stage (its leaving
            // row is one of my rows / its entering column one of my warp's): rare, so a compact rolled loop
            // that takes r_u from the chunk's shared-memory slice instead of the register copy — same values,
            // same operations in the same order
            for (int u = 0; u < t; u++) {
              double av[kSwR];
              load_a(u, av);
              double rv[kC];
#pragma unroll
              for (int h = 0; h < kH; h++) {
                const double2 v = act[h] ? *reinterpret_cast<const double2*>(srb + (size_t)u * bw + jt + 64 * h)
                                         : make_double2(0.0, 0.0);
                rv[2 * h] = v.x;
                rv[2 * h + 1] = v.y;
              }
              const bool hit = (smask >> u) & 1u;
              const int lk = hit ? s_l[u] - i : -1;      // its leaving row among my rows (else out of 0..kSwR-1)
              const int ce = hit ? s_e[u] - j : -1;      // its entering column among mine: 0, 1 (, 64, 65)
              const double pu = s_p[u];
#pragma unroll
              for (int k = 0; k < kSwR; k++) {
#pragma unroll
                for (int h = 0; h < kH; h++) {
                  if (k == lk) {                         // LPState.java:137-146
                    x[k][h].x = rv[2 * h];
                    x[k][h].y = rv[2 * h + 1];
                  } else {
                    const double q = (ce == 64 * h || ce == 64 * h + 1) ? -ddiv_call(av[k], pu) : 0.0;   // :157 / :172
                    x[k][h].x = (ce == 64 * h) ? q : __dsub_rn(x[k][h].x, __dmul_rn(av[k], rv[2 * h]));  // :162-164 / :177
                    x[k][h].y = (ce == 64 * h + 1) ? q : __dsub_rn(x[k][h].y, __dmul_rn(av[k], rv[2 * h + 1]));
                  }
                }
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage]);     // the tile and its a-slice are consumed
          if (++stage == kSwStages) { stage = 0; phase ^= 1u; }
          if (nst > 1) ready = mbar_try_wait(&full[stage], phase);   // the probe overlaps the stores
#pragma unroll
          for (int k = 0; k < kSwR; k++)
            if (i + k < a.rows) {
#pragma unroll
              for (int h = 0; h < kH; h++)
                if (act[h]) st128_stream(out + (long long)k * a.ld + 64 * h, x[k][h].x, x[k][h].y);
            }
          i += kSwSR;
          out += (long long)kSwSR * a.ld;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&r_empty[rb]);        // the slice (and its descriptor slot) may be refilled
      }
    }
  }
  // last pass CTA retires the block
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned int tk = atomicAdd(&ctl->blk_ticket, 1u);
    s_last = (tk == (unsigned int)a.ncta - 1);
  }
  __syncthreads();
  if (s_last && tid == 0) {
    ctl->blk_ticket = 0;
    ctl->blk_queue = 0;
    ctl->blk_pend[set] = 0;
    // nobody in THIS launch reads cur_at[q ^ 1]: the next launch does
    ctl->cur_at[a.q ^ 1] = (!a.inplace && t > 0) ? (cur ^ 1) : cur;
    if (t > 0) ctl->sweeps_done += 1;
    __threadfence();
  }
}

// the pass as a kernel of its own (every CTA of the grid)
template <class Shape>
__global__ void __launch_bounds__(Shape::kThreads, 1)
kb_sweep(const __grid_constant__ SweepArgs a, const __grid_constant__ CUtensorMap tmT0,
         const __grid_constant__ CUtensorMap tmT1, const __grid_constant__ CUtensorMap tmA,
         const __grid_constant__ CUtensorMap tmR) {
  extern __shared__ __align__(1024) unsigned char sweep_smem[];
  sweep_role<Shape>(a, &tmT0, &tmT1, &tmA, &tmR, sweep_smem);
}

}  // namespace lps
