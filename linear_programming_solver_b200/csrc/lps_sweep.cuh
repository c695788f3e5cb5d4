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

constexpr unsigned long long kSwWaitNs = 4ull * 1000ull * 1000ull * 1000ull;

// Shape of the pass.
//   kS    most pending pivots: their rows r_u live in the consumer threads' registers (2 kS kC of them)
//   kR    rows per consumer thread per sub-pass;  kC  columns per thread (2, or 4 as two pairs 64 apart)
//   kCW   consumer warps;  kColWarps of them side by side span the strip (kColWarps x 32 kC columns, at most
//         256: one TMA box), the kCW / kColWarps "row lanes" are stacked over the rows
//   kP    sub-passes per stage: a stage holds kP x (row lanes x kR) rows, and a warp works through its kP row
//         groups one after the other under ONE full / empty barrier round trip
// What drives the choice (ncu, profiles/r02_pass_shapes.md): every FP64 instruction needs one a_u[i] operand,
// delivered to every lane through the shared-memory pipe and reused by the thread's kC columns; the per-stage
// bookkeeping (barrier probe, tile load, stores) is a latency chain that only OTHER warps' arithmetic hides,
// so the FP64 pipe wants three or four consumer warps per scheduler; and registers come in units of four
// warps per CTA (12 warps: 168 per thread, 16 warps: 128) while spills are expensive here — next to 220 KB of
// shared memory the L1 that would cache them is almost gone.
//   kPipe software-pipelined consumer (kP == 1): the tile of stage s+1 is fetched into a second register set in
//         the middle of stage s's arithmetic (its barrier probe is issued before the arithmetic starts), so the
//         barrier / shared-memory latencies of the per-stage bookkeeping hide under the warp's OWN FP64 work
template <int kS_, int kR_, int kC_, int kCW_, int kColWarps_, int kP_, bool kPipe_ = false>
struct SweepShape {
  static constexpr int kS = kS_, kR = kR_, kC = kC_, kCW = kCW_, kColWarps = kColWarps_, kP = kP_;
  static constexpr bool kPipe = kPipe_;
  static_assert(!kPipe || kP == 1, "the pipelined consumer handles one sub-pass per stage");
  static constexpr int kCols = kColWarps * 32 * kC;          // strip width
  static constexpr int kRowLanes = kCW / kColWarps;
  static constexpr int kGR = kR * kRowLanes;                 // rows per sub-pass
  static constexpr int kSR = kGR * kP;                       // rows per stage
  static constexpr int kThreads = (kCW + 1 + 3) / 4 * 4 * 32;   // + the producer warp, in whole units of four warps
  static constexpr size_t kTile = (size_t)kSR * kCols * sizeof(double);
  static constexpr size_t kASlice = ((size_t)kS * kSR * sizeof(double) + 127) / 128 * 128;
  static constexpr size_t kRSlice = (size_t)kS * kCols * sizeof(double);
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
  static_assert((kC == 2 || kC == 4) && kCW % kColWarps == 0 && kR % 2 == 0 && kCols <= 256, "shape");
  static_assert(kStages >= 3, "pipeline depth");
  static_assert(kTile % 128 == 0 && kRSlice % 128 == 0 && (kSR * sizeof(double)) % 16 == 0, "TMA alignment");
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
// non-blocking probe (no suspend): has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test_wait(unsigned long long* bar, unsigned int parity) {
  unsigned int ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
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

// The rare path of the pass, out of line so that its registers (and its division calls) do not weigh on the
// allocation of the hot loop: a partial block, or a pending pivot OVERWRITES some of the warp's cells in this
// sub-pass (its leaving row is one of the thread's rows / its entering column one of the warp's).  A compact
// rolled loop that takes r_u from the chunk's shared-memory slice instead of the register copy — same values,
// same operations in the same order.  The thread's cells travel by value (in registers, both ways).
template <int kR, int kH>
struct SweepCells {
  double2 v[kR][kH];
};
template <int kR, int kH>
__device__ __noinline__ SweepCells<kR, kH> sweep_special(SweepCells<kR, kH> x, unsigned int sa /* shared address of a_0[i] */,
                                                         int a_pitch /* kSR */, const double* srb /* r slice + my column */,
                                                         int bw, int t, unsigned int smask, int i, int j,
                                                         unsigned int actbits, const int* s_l, const int* s_e,
                                                         const double* s_p) {
  for (int u = 0; u < t; u++) {
    double av[kR];
#pragma unroll
    for (int k = 0; k < kR; k += 2) {
      const double2 v = lds128(sa + (unsigned int)((u * a_pitch + k) * sizeof(double)));
      av[k] = v.x;
      av[k + 1] = v.y;
    }
    double rv[2 * kH];
#pragma unroll
    for (int h = 0; h < kH; h++) {
      const double2 v = ((actbits >> h) & 1u) ? *reinterpret_cast<const double2*>(srb + (size_t)u * bw + 64 * h)
                                              : make_double2(0.0, 0.0);
      rv[2 * h] = v.x;
      rv[2 * h + 1] = v.y;
    }
    const bool hit = (smask >> u) & 1u;
    const int lk = hit ? s_l[u] - i : -1;      // its leaving row among my rows (else out of 0..kR-1)
    const int ce = hit ? s_e[u] - j : -1;      // its entering column among mine: 0, 1 (, 64, 65)
    const double pu = s_p[u];
#pragma unroll
    for (int k = 0; k < kR; k++) {
#pragma unroll
      for (int h = 0; h < kH; h++) {
        if (k == lk) {                         // LPState.java:137-146
          x.v[k][h].x = rv[2 * h];
          x.v[k][h].y = rv[2 * h + 1];
        } else {
          const double q = (ce == 64 * h || ce == 64 * h + 1) ? -ddiv_call(av[k], pu) : 0.0;   // :157 / :172
          x.v[k][h].x = (ce == 64 * h) ? q : __dsub_rn(x.v[k][h].x, __dmul_rn(av[k], rv[2 * h]));      // :162-164 / :177
          x.v[k][h].y = (ce == 64 * h + 1) ? q : __dsub_rn(x.v[k][h].y, __dmul_rn(av[k], rv[2 * h + 1]));
        }
      }
    }
  }
  return x;
}

struct SweepArgs {
  CtlS* ctl;
  double* Tbuf[2];             // the tableau buffers, (mloc+1) x ld each; ctl->cur_at[q] names the current one.
                               // In place: both entries are the same buffer.
  long long ld;
  int rows;                    // mloc + 1 (objective row included)
  int chunk_rows;              // rows per chunk, a multiple of the stage height
  int bw;                      // tile width in columns = min(Shape::kCols, ld): box of the T and pending-row maps
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
  constexpr int kS = Shape::kS, kR = Shape::kR, kC = Shape::kC, kH = kC / 2, kP = Shape::kP;
  constexpr int kGR = Shape::kGR, kSR = Shape::kSR, kStages = Shape::kStages, kCW = Shape::kCW;
  constexpr int kCols = Shape::kCols;
  CtlS* const ctl = a.ctl;
  const int set = a.q;
  const int t = ctl->blk_pend[set];
  const int cur = ctl->cur_at[a.q];
  const CUtensorMap* const tmT = cur ? tmT1 : tmT0;
  double* const dst = a.Tbuf[a.inplace ? cur : (cur ^ 1)];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ unsigned long long s_t0;
  if (tid == 0) s_t0 = globaltimer_ns();
  double* const tiles = reinterpret_cast<double*>(smem + SM::kOffTiles);
  double* const s_a = reinterpret_cast<double*>(smem + SM::kOffA);
  double* const s_r = reinterpret_cast<double*>(smem + SM::kOffR);
  unsigned long long* const full = reinterpret_cast<unsigned long long*>(smem + SM::kOffBars);
  unsigned long long* const empty = full + kStages;
  unsigned long long* const r_full = empty + kStages;       // [2]
  unsigned long long* const r_empty = r_full + 2;            // [2]
  int4* const desc = reinterpret_cast<int4*>(smem + SM::kOffMeta);      // [2] chunk descriptor {j0, first row, stages, stop}
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
      for (int s = 0; s < kStages; s++) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], kCW);
      }
      for (int b2 = 0; b2 < 2; b2++) {
        mbar_init(&r_full[b2], 1);
        mbar_init(&r_empty[b2], kCW);
      }
      mbar_fence_init();
    }
    __syncthreads();

    const int bw = a.bw;
    const int nstrips = (int)((a.ld + bw - 1) / bw);
    const int nbands = (a.rows + a.chunk_rows - 1) / a.chunk_rows;
    const long long nchunks = (long long)nstrips * nbands;
    const unsigned int tile_bytes = (unsigned int)(kSR * bw * sizeof(double));
    const unsigned int a_bytes = (unsigned int)(a.bu * kSR * sizeof(double));
    const unsigned int r_bytes = (unsigned int)(a.bu * bw * sizeof(double));

    if (warp > kCW) {
      // padding warps of the register-allocation unit: nothing to do in the pass
    } else if (warp == kCW) {
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
          desc[rb] = make_int4(j0, ib, (iend - ib + kSR - 1) / kSR, 0);
          mbar_arrive_expect_tx(&r_full[rb], r_bytes);
          tma_load_2d(s_r + (size_t)rb * kS * kCols, tmR, j0, 0, &r_full[rb]);
          for (int i0 = ib; i0 < iend; i0 += kSR) {
            mbar_wait(&empty[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&full[stage], tile_bytes + a_bytes);
            tma_load_2d_hint(tiles + (size_t)stage * kSR * kCols, tmT, j0, i0, &full[stage], pol);
            tma_load_2d(smem + SM::kOffA + (size_t)stage * SM::kASlice, tmA, i0, 0, &full[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    } else {
      // ---------------- consumers ----------------
      const int cgrp = warp % Shape::kColWarps, rlane = warp / Shape::kColWarps;
      const int wcol = cgrp * 32 * kC;                     // first column of my warp inside the strip
      const int jt = wcol + lane * 2;                      // my first pair; the second one (kC == 4) is 64 further
      // byte offsets of my cells inside a stage's tile / a-slice: fixed for the whole pass
      const unsigned int row_b = (unsigned int)(bw * sizeof(double));
      const unsigned int tiles_s = smem_u32(tiles) + (unsigned int)(rlane * kR) * row_b + (unsigned int)(jt * sizeof(double));
      const unsigned int sa_s = smem_u32(s_a) + (unsigned int)(rlane * kR * sizeof(double));
      double r[kS][kC];
#pragma unroll
      for (int u = 0; u < kS; u++)
#pragma unroll
        for (int c = 0; c < kC; c++) r[u][c] = 0.0;
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
        const double* const srb = s_r + (size_t)rb * kS * kCols;    // stays valid until I release it below
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
        unsigned int colmask = 0;                        // pending pivots whose entering column my WARP holds
        unsigned int rowmask = 0;                        // pending pivots whose leaving row lies in this chunk
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
        unsigned int actbits = 0;
#pragma unroll
        for (int h = 0; h < kH; h++) actbits |= act[h] ? (1u << h) : 0u;
        int i = ib + rlane * kR;                         // my first row of the stage's first sub-pass
        double* out = dst + (long long)i * a.ld + j;
        if constexpr (Shape::kPipe) {
          // ---- the stages of the chunk, software-pipelined over two register sets ----
          auto load_x = [&](double2 (&x)[kR][kH], int st) {
            const unsigned int tile = tiles_s + (unsigned int)st * (unsigned int)SM::kTile;
#pragma unroll
            for (int k = 0; k < kR; k++)
#pragma unroll
              for (int h = 0; h < kH; h++) x[k][h] = lds128(tile + k * row_b + 64 * h * (unsigned int)sizeof(double));
          };
          auto step = [&](double2 (&x)[kR][kH], double2 (&xn)[kR][kH], bool has_next) {
            const int cs = stage;                          // the stage this step consumes
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
            const bool nready = has_next ? mbar_test_wait(&full[stage], phase) : true;   // answer needed mid-arithmetic
            const unsigned int sa = sa_s + (unsigned int)cs * (unsigned int)SM::kASlice;
            unsigned int smask = colmask;
            if (rowmask != 0) {
              for (int u = 0; u < t; u++) {
                const int d = s_l[u] - i;
                if (d >= 0 && d < kR) smask |= 1u << u;
              }
            }
            auto load_a = [&](int u, double (&av)[kR]) {
#pragma unroll
              for (int k = 0; k < kR; k += 2) {
                const double2 v = lds128(sa + (unsigned int)((u * kSR + k) * sizeof(double)));
                av[k] = v.x;
                av[k + 1] = v.y;
              }
            };
            auto update = [&](int u, const double (&av)[kR]) {
#pragma unroll
              for (int k = 0; k < kR; k++)
#pragma unroll
                for (int h = 0; h < kH; h++) {
                  x[k][h].x = __dsub_rn(x[k][h].x, __dmul_rn(av[k], r[u][2 * h]));      // LPState.java:162-164 / :177
                  x[k][h].y = __dsub_rn(x[k][h].y, __dmul_rn(av[k], r[u][2 * h + 1]));
                }
            };
            auto fetch_next = [&]() {
              if (has_next) {
                if (!nready) mbar_wait(&full[stage], phase);
                load_x(xn, stage);
              }
            };
            if (smask == 0 && t == kS) {
#pragma unroll
              for (int u = 0; u < kS / 2; u++) {
                double av[kR];
                load_a(u, av);
                update(u, av);
              }
              fetch_next();                                // lands while the second half is computed
#pragma unroll
              for (int u = kS / 2; u < kS; u++) {
                double av[kR];
                load_a(u, av);
                update(u, av);
              }
            } else {
              SweepCells<kR, kH> xs;
#pragma unroll
              for (int k = 0; k < kR; k++)
#pragma unroll
                for (int h = 0; h < kH; h++) xs.v[k][h] = x[k][h];
              xs = sweep_special<kR, kH>(xs, sa, kSR, srb + jt, bw, t, smask, i, j, actbits, s_l, s_e, s_p);
#pragma unroll
              for (int k = 0; k < kR; k++)
#pragma unroll
                for (int h = 0; h < kH; h++) x[k][h] = xs.v[k][h];
              fetch_next();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[cs]);        // the a-slice of the stage is consumed (the tile long since)
#pragma unroll
            for (int k = 0; k < kR; k++)
              if (i + k < a.rows) {
#pragma unroll
                for (int h = 0; h < kH; h++)
                  if (act[h]) st128_stream(out + (long long)k * a.ld + 64 * h, x[k][h].x, x[k][h].y);
              }
            i += kSR;
            out += (long long)kSR * a.ld;
          };
          double2 xa[kR][kH], xb[kR][kH];
          mbar_wait(&full[stage], phase);
          load_x(xa, stage);
          while (nst > 0) {
            step(xa, xb, nst > 1);
            if (--nst == 0) break;
            step(xb, xa, nst > 1);
            --nst;
          }
        } else {
        // ---- the stages of the chunk ----
        bool ready = mbar_try_wait(&full[stage], phase);
        for (; nst > 0; nst--) {
          if (!ready) mbar_wait(&full[stage], phase);
          const unsigned int tile0 = tiles_s + (unsigned int)stage * (unsigned int)SM::kTile;
          const unsigned int sa0 = sa_s + (unsigned int)stage * (unsigned int)SM::kASlice;
#pragma unroll
          for (int p = 0; p < kP; p++) {
            // my kR rows of sub-pass p: rows i + p kGR .. of the tableau, rows (p kRowLanes + rlane) kR .. of the tile
            const int ip = i + p * kGR;
            const unsigned int tile = tile0 + (unsigned int)(p * kGR) * row_b;
            const unsigned int sa = sa0 + (unsigned int)(p * kGR * sizeof(double));
            unsigned int smask = colmask;                // ... or whose leaving row is one of my rows
            if (rowmask != 0) {
              for (int u = 0; u < t; u++) {
                const int d = s_l[u] - ip;
                if (d >= 0 && d < kR) smask |= 1u << u;
              }
            }
            double2 x[kR][kH];
#pragma unroll
            for (int k = 0; k < kR; k++)
#pragma unroll
              for (int h = 0; h < kH; h++) x[k][h] = lds128(tile + k * row_b + 64 * h * (unsigned int)sizeof(double));
            // one pending pivot on my kR x kC cells: a_u[ip..ip+kR) from shared memory (warp-wide broadcasts)
            auto load_a = [&](int u, double (&av)[kR]) {
#pragma unroll
              for (int k = 0; k < kR; k += 2) {
                const double2 v = lds128(sa + (unsigned int)((u * kSR + k) * sizeof(double)));
                av[k] = v.x;
                av[k + 1] = v.y;
              }
            };
            auto update = [&](int u, const double (&av)[kR]) {
#pragma unroll
              for (int k = 0; k < kR; k++)
#pragma unroll
                for (int h = 0; h < kH; h++) {
                  x[k][h].x = __dsub_rn(x[k][h].x, __dmul_rn(av[k], r[u][2 * h]));      // LPState.java:162-164 / :177
                  x[k][h].y = __dsub_rn(x[k][h].y, __dmul_rn(av[k], r[u][2 * h + 1]));
                }
            };
            if (smask == 0 && t == kS) {
              // the common case, free of branches
#pragma unroll
              for (int u = 0; u < kS; u++) {
                double av[kR];
                load_a(u, av);
                update(u, av);
              }
            } else {
              SweepCells<kR, kH> xs;
#pragma unroll
              for (int k = 0; k < kR; k++)
#pragma unroll
                for (int h = 0; h < kH; h++) xs.v[k][h] = x[k][h];
              xs = sweep_special<kR, kH>(xs, sa, kSR, srb + jt, bw, t, smask, ip, j, actbits, s_l, s_e, s_p);
#pragma unroll
              for (int k = 0; k < kR; k++)
#pragma unroll
                for (int h = 0; h < kH; h++) x[k][h] = xs.v[k][h];
            }
            if (p == kP - 1) {
              __syncwarp();
              if (lane == 0) mbar_arrive(&empty[stage]);   // the tile and its a-slice are consumed
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
              if (nst > 1) ready = mbar_try_wait(&full[stage], phase);   // the probe overlaps the stores
            }
#pragma unroll
            for (int k = 0; k < kR; k++)
              if (ip + k < a.rows) {
#pragma unroll
                for (int h = 0; h < kH; h++)
                  if (act[h]) st128_stream(out + (long long)(p * kGR + k) * a.ld + 64 * h, x[k][h].x, x[k][h].y);
              }
          }
          i += kSR;
          out += (long long)kSR * a.ld;
        }
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
    if (t > 0) {
      ctl->sweeps_done += 1;
      ctl->dbg_ns[10] += globaltimer_ns() - s_t0;      // the pass's own clock: the last CTA to retire (host: split tuning)
      ctl->dbg_ns[11] += 1;
    }
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
