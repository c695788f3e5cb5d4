// C-ABI implementation (include/lps_b200.h) over the kernels in lps_kernels.cuh.
// Host responsibilities only: allocation, host<->HBM copies, launch sequencing in batches,
// polling of the device control block.  No arithmetic on tableau values happens on the host
// and there is no CPU fallback: without a CUDA device every entry point fails.
#include "../../include/lps_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "lps_kernels.cuh"
#include "lps_sharded.cuh"
#include "lps_loop.cuh"
#include "lps_blocked.cuh"
#include "lps_sweep.cuh"
#include "lps_step.cuh"

using namespace lps;

// layouts the ctypes / JNI bindings rely on (tests/test_abi.py checks the Python side)
static_assert(sizeof(lps_options) == 64, "lps_options layout");
static_assert(sizeof(lps_run_result) == 64, "lps_run_result layout");
static_assert(sizeof(lps_objective_op) == 16, "lps_objective_op layout");

struct lps_handle_s {
  lps_options opt;
  int dev = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 0;

  bool loaded = false;
  int m = 0, n = 0;
  long long ld = 0;
  double* T = nullptr;
  size_t T_bytes = 0;
  double *col0 = nullptr, *col1 = nullptr, *bcol = nullptr, *rowbuf = nullptr, *scratch = nullptr;
  size_t vec_cap = 0;  // capacity (doubles) of col0/col1/bcol/scratch; rowbuf has its own
  size_t row_cap = 0;
  int* pos2var = nullptr;
  size_t pos_cap = 0;
  int2* plog = nullptr;
  long long log_cap = 1ll << 22;
  CtlS* ctls = nullptr;   // device control block (single-GPU kernels use its `base`)
  Ctl* ctl = nullptr;     // == &ctls->base
  CtlS* h_ctls = nullptr; // pinned host copy
  Ctl* h_ctl = nullptr;   // == &h_ctls->base
  Cand* partials = nullptr;
  ObjOp* d_ops = nullptr;
  size_t ops_cap = 0;

  // device-side (e_next, colbuf[npivots&1], bcol) are consistent with the tableau
  bool next_valid = false;
  // where the staged entering index lives: false = Ctl::e_next (three-kernel single-GPU path),
  // true = CtlS::e_nx[(npivots+1)&1] (sharded kernels and the persistent loop)
  bool next_in_nx = false;
  int loop_grid = 0;  // co-resident CTAs of k_loop (0 = not queried yet)
  // colbuf[npivots&1] holds this column (or -1)
  int col_holds = -1;
  long long total_pivots = 0;

  // row-sharded mode (lps_shard_*): this handle holds rows [row0,row1) of an m_total-row LP
  bool sharded = false;
  int rank = 0, world = 1, m_total = 0, row0 = 0, row1 = 0;
  CommBlock* comm = nullptr;
  size_t comm_bytes = 0;
  Peers peers{};
  bool attached = false;
  std::vector<void*> ipc_opened;

  // blocked loop (lps_blocked.cuh): up to `block` pivots deferred between two tableau passes
  int block = 1;            // resolved from opt.block_pivots at create
  double* acols = nullptr;  // [block][apitch] pending entering columns
  long long apitch = 0;
  size_t acols_cap = 0;     // doubles
  int flush_grid = 0;       // persistent grid of kb_flush (0 = not sized yet)
  int panel_grid = 0;       // cooperative grid of kb_panel (0 = not sized yet, -1 = unavailable)
  PeerCand* ppartials = nullptr;          // kb_panel: tagged ratio-test partials, one per CTA
  unsigned long long* pmins = nullptr;    // kb_panel: per-CTA minima
  unsigned int* psync = nullptr;          // kb_panel: ticket / go words of its two grid-wide syncs
  unsigned int panel_launches = 0;        // tag source: never reset, so a stale slot can never match
  bool panel_dirty = false;               // a run was abandoned inside a grid sync: re-arm the sync words
  long long ll_off = 0;                   // byte offset of the packet area inside the exchange block

  // TMA pass (lps_sweep.cuh) and look-ahead loop (lps_step.cuh)
  double* T2 = nullptr;                   // second tableau buffer of the out-of-place pass (look-ahead loop only)
  size_t T2_bytes = 0;
  double *bvec = nullptr, *cvec = nullptr;   // running b column / objective row of the look-ahead panel
  size_t bvec_cap = 0, cvec_cap = 0;
  CUtensorMap tm_T[2];                    // tableau buffers (tm_T[1] == tm_T[0] when the pass runs in place)
  CUtensorMap tm_A[2], tm_R[2];           // pending columns / rows of set 0 and 1
  bool tm_valid = false;
  int tm_rows_set = 0;                    // pending sets the maps were built for (1: in place, 2: look-ahead)
  int step_grid = 0;                      // cooperative grid of kb_step (0 = not sized yet)
  int sweep_grid = 0;                     // grid of the stand-alone kb_sweep (0 = not sized yet)
  unsigned int look_launches = 0;         // launch counter of the look-ahead loop: parity + tag source
  int tuned_P = 0;                        // panel CTAs chosen by tune_split() for a tableau of tuned_m x tuned_ld
  int tuned_m = 0;
  long long tuned_ld = 0;

  std::vector<cudaEvent_t> ev;  // time_kernels event pool
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  std::string err;
};

namespace {

constexpr int kRatioThreads = 256;
constexpr int kDefaultBlock = 16;   // pivots per tableau pass of the blocked loop
constexpr int kFlushMaxBlock = kPanelMax;  // kb_flush's double-buffered operand slices must fit in 227 KB of shared memory

int fail(lps_handle h, int code, const char* what, cudaError_t ce = cudaSuccess) {
  if (h) {
    h->err = what;
    if (ce != cudaSuccess) {
      h->err += ": ";
      h->err += cudaGetErrorString(ce);
    }
  }
  return code;
}

#define CK(call)                                                     \
  do {                                                               \
    cudaError_t ce_ = (call);                                        \
    if (ce_ != cudaSuccess) return fail(h, LPS_ERR_CUDA, #call, ce_); \
  } while (0)

inline long long round_up(long long x, long long a) { return (x + a - 1) / a * a; }
inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

int free_tableau(lps_handle h) {
  if (h->T) cudaFree(h->T);
  h->T = nullptr;
  h->T_bytes = 0;
  return 0;
}

int ensure_buffers(lps_handle h, int m, int n_cols /* n incl. any aux column */) {
  long long ld = round_up((long long)n_cols + 1, 16);
  size_t need = (size_t)(m + 1) * (size_t)ld * sizeof(double);
  if (need > h->T_bytes) {
    free_tableau(h);
    cudaError_t ce = cudaMalloc(&h->T, need);
    if (ce != cudaSuccess) return fail(h, LPS_ERR_NOMEM, "cudaMalloc(tableau)", ce);
    h->T_bytes = need;
  }
  size_t vneed = (size_t)m + 1 + 16;
  if (vneed > h->vec_cap) {
    for (double** p : {&h->col0, &h->col1, &h->bcol, &h->scratch}) {
      if (*p) cudaFree(*p);
      cudaError_t ce = cudaMalloc(p, vneed * sizeof(double));
      if (ce != cudaSuccess) return fail(h, LPS_ERR_NOMEM, "cudaMalloc(column staging)", ce);
    }
    h->vec_cap = vneed;
  }
  if ((size_t)ld > h->row_cap) {
    if (h->rowbuf) cudaFree(h->rowbuf);
    cudaError_t ce = cudaMalloc(&h->rowbuf, (size_t)ld * sizeof(double));
    if (ce != cudaSuccess) return fail(h, LPS_ERR_NOMEM, "cudaMalloc(row staging)", ce);
    h->row_cap = (size_t)ld;
  }
  size_t pneed = (size_t)m + n_cols + 16;
  if (pneed > h->pos_cap) {
    if (h->pos2var) cudaFree(h->pos2var);
    cudaError_t ce = cudaMalloc(&h->pos2var, pneed * sizeof(int));
    if (ce != cudaSuccess) return fail(h, LPS_ERR_NOMEM, "cudaMalloc(positions)", ce);
    h->pos_cap = pneed;
  }
  h->m = m;
  h->n = n_cols;
  h->ld = ld;
  h->tm_valid = false;   // tensor maps of the TMA pass describe the old buffers / shape
  if (h->block > 1) {
    h->apitch = round_up((long long)m + 1 + 512, 8);   // kb_flush copies whole chunks of rows
    size_t aneed = (size_t)2 * h->block * (size_t)h->apitch + 512;   // two pending sets (look-ahead loop)
    if (aneed > h->acols_cap) {
      if (h->acols) cudaFree(h->acols);
      h->acols = nullptr;
      cudaError_t ce = cudaMalloc(&h->acols, aneed * sizeof(double));
      if (ce != cudaSuccess) return fail(h, LPS_ERR_NOMEM, "cudaMalloc(pending columns)", ce);
      h->acols_cap = aneed;
      cudaMemsetAsync(h->acols, 0, aneed * sizeof(double), h->stream);
    }
  }
  return LPS_OK;
}

int ensure_comm(lps_handle h);

// a single-GPU handle runs the blocked kernels as a world of one talking to itself
int ensure_self_comm(lps_handle h) {
  if (h->block <= 1 || h->sharded) return LPS_OK;
  if (h->ld > (long long)kMaxChunks * kChunk) {   // too wide for the flag slots: unblocked
    // a comm block left over from an earlier, narrower load would make use_blocked() say yes and the
    // row kernels index their flag slots out of bounds
    if (h->comm && !h->attached) {
      cudaFree(h->comm);
      h->comm = nullptr;
      h->comm_bytes = 0;
    }
    return LPS_OK;
  }
  h->rank = 0;
  h->world = 1;
  h->m_total = h->m;
  h->row0 = 0;
  h->row1 = h->m;
  return ensure_comm(h);
}

int reset_state(lps_handle h) {
  CK(cudaMemsetAsync(h->ctls, 0, sizeof(CtlS), h->stream));
  int tot = h->m + h->n;
  k_iota<<<cdiv(tot, 256), 256, 0, h->stream>>>(h->pos2var, tot);
  CK(cudaGetLastError());
  h->next_valid = false;
  h->col_holds = -1;
  h->total_pivots = 0;
  h->loaded = true;
  return LPS_OK;
}

int sync_ctl(lps_handle h) {
  CK(cudaMemcpyAsync(h->h_ctls, h->ctls, sizeof(CtlS), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return LPS_OK;
}

int ratio_grid(lps_handle h) {
  return std::max(1, std::min(cdiv(h->m, kRatioThreads), 2 * h->sm_count));
}

// launch K2 + K3 for the pivot already committed in ctl
int launch_scale_update(lps_handle h, cudaEvent_t e0, cudaEvent_t e1) {
  k_scale_row<<<cdiv(h->ld, 256), 256, 0, h->stream>>>(h->ctl, h->T, h->ld, h->m, h->n, h->rowbuf,
                                                      h->col0, h->col1, h->opt.epsilon);
  if (e0) cudaEventRecord(e0, h->stream);
#define LPS_UPD(T_, R_, U_, B_)                                                               \
  do {                                                                                        \
    dim3 grid(cdiv(h->ld, 4ll * (T_)), cdiv(h->m + 1, (R_)));                                 \
    k_update<T_, R_, U_, B_><<<grid, (T_), 0, h->stream>>>(h->ctl, h->T, h->ld, h->m, h->n,   \
                                                          h->rowbuf, h->col0, h->col1, h->bcol); \
  } while (0)
  switch (h->opt.update_variant) {
    case 0: LPS_UPD(256, 32, 8, 1); break;
    case 1: LPS_UPD(256, 32, 4, 2); break;
    case 2: LPS_UPD(256, 64, 8, 1); break;
    default:  // best of the first B200 sweep (profiles/r01_update_variants.md)
    case 3: LPS_UPD(128, 32, 8, 2); break;
    case 4: LPS_UPD(512, 32, 4, 1); break;
    case 5: LPS_UPD(256, 32, 8, 2); break;
    case 6: LPS_UPD(128, 64, 4, 4); break;
    case 7: LPS_UPD(256, 16, 4, 3); break;
  }
#undef LPS_UPD
  if (e1) cudaEventRecord(e1, h->stream);
  return LPS_OK;
}

int launch_ratio(lps_handle h, int mode) {
  k_ratio<<<ratio_grid(h), kRatioThreads, 0, h->stream>>>(h->ctl, h->col0, h->col1, h->bcol, h->m,
                                                         h->n, h->opt.epsilon, h->opt.inf,
                                                         h->partials, h->plog, h->log_cap,
                                                         h->pos2var, mode);
  return LPS_OK;
}

// make (e_next, colbuf, bcol) valid on the device
int prepare_next(lps_handle h) {
  if (h->next_valid) return LPS_OK;
  k_begin_run<<<1, 1, 0, h->stream>>>(h->ctl, -1, 1);
  k_first_positive<<<std::max(1, cdiv(h->n, 256)), 256, 0, h->stream>>>(h->ctl, h->T + (long long)h->m * h->ld,
                                                                       h->n, h->opt.epsilon);
  k_extract<<<cdiv(h->m + 1, 256), 256, 0, h->stream>>>(h->ctl, h->T, h->ld, h->m, h->n, -1, h->col0,
                                                       h->col1, h->bcol);
  CK(cudaGetLastError());
  h->next_valid = true;
  h->next_in_nx = false;
  h->col_holds = -2;  // "whatever e_next is"
  return LPS_OK;
}


// ---- row-sharded mode ---------------------------------------------------------------------
int ensure_comm(lps_handle h) {
  // pivot-row store: 2 parity slots for the pivot-per-pass kernels, 2 sets of `block` slots for the blocked loops
  // ... then the packet area of kb_panel's exchange: LLPacket row[2][ld], LLPacket cand[2][kMaxRanks][4].
  // Every rank of a sharded solve must be created with the same block_pivots: the offsets are shared.
  h->ll_off = (long long)round_up((long long)(sizeof(CommBlock) + (size_t)std::max(2, 2 * h->block) * (size_t)h->ld * sizeof(double)), 16);
  size_t need = (size_t)h->ll_off + (2 * (size_t)h->ld + 2 * kMaxRanks * 4) * sizeof(LLPacket);
  if (need > h->comm_bytes) {
    if (h->attached) return fail(h, LPS_ERR_STATE, "shard: tableau grew after peers were attached");
    if (h->comm) cudaFree(h->comm);
    cudaError_t ce = cudaMalloc(&h->comm, need);
    if (ce != cudaSuccess) return fail(h, LPS_ERR_NOMEM, "cudaMalloc(comm block)", ce);
    h->comm_bytes = need;
  }
  CK(cudaMemsetAsync(h->comm, 0, h->comm_bytes, h->stream));
  // until peers are attached the rank talks to itself (world == 1 works out of the box)
  for (int k = 0; k < kMaxRanks; k++) {
    h->peers.blk[k] = h->comm;
    h->peers.rowbuf[k] = reinterpret_cast<double*>(reinterpret_cast<char*>(h->comm) + sizeof(CommBlock));
  }
  return LPS_OK;
}

int shard_setup(lps_handle h, int m_total, int n, int rank, int world) {
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world || m_total < 0 || n < 0)
    return fail(h, LPS_ERR_INVALID, "shard: bad rank/world/dimensions");
  // row_flag has kMaxChunks slots per parity; the persistent loop uses one per kLoopChunk (128) columns
  if ((long long)n + 1 > (long long)kMaxChunks * std::min(kChunk, kLoopChunk))
    return fail(h, LPS_ERR_INVALID, "shard: too many columns");
  h->sharded = true;
  h->rank = rank;
  h->world = world;
  h->m_total = m_total;
  h->row0 = (int)(((long long)rank * m_total) / world);        // LPState.java:222
  h->row1 = (int)(((long long)(rank + 1) * m_total) / world);  // LPState.java:223
  int rc = ensure_buffers(h, h->row1 - h->row0, n);
  if (rc) return rc;
  // positions are global: n + m_total entries
  size_t pneed = (size_t)m_total + n + 16;
  if (pneed > h->pos_cap) {
    if (h->pos2var) cudaFree(h->pos2var);
    cudaError_t ce = cudaMalloc(&h->pos2var, pneed * sizeof(int));
    if (ce != cudaSuccess) return fail(h, LPS_ERR_NOMEM, "cudaMalloc(positions)", ce);
    h->pos_cap = pneed;
  }
  return ensure_comm(h);
}

int shard_reset_state(lps_handle h) {
  CK(cudaMemsetAsync(h->ctls, 0, sizeof(CtlS), h->stream));
  int tot = h->m_total + h->n;
  k_iota<<<cdiv(tot, 256), 256, 0, h->stream>>>(h->pos2var, tot);
  CK(cudaGetLastError());
  h->next_valid = false;
  h->col_holds = -1;
  h->total_pivots = 0;
  h->loaded = true;
  return LPS_OK;
}

int shard_prepare_next(lps_handle h) {
  if (h->next_valid) return LPS_OK;
  ks_begin_run<<<1, 1, 0, h->stream>>>(h->ctls, -1, 1);
  ks_first_positive<<<std::max(1, cdiv(h->n, 256)), 256, 0, h->stream>>>(h->ctls, h->T + (long long)h->m * h->ld, h->n,
                                                                        h->opt.epsilon);
  ks_extract<<<cdiv(h->m + 1, 256), 256, 0, h->stream>>>(h->ctls, h->T, h->ld, h->m, h->n, -1, h->col0,
                                                        h->col1, h->bcol);
  CK(cudaGetLastError());
  h->next_valid = true;
  h->next_in_nx = true;
  return LPS_OK;
}

void shard_launch_pivot(lps_handle h, bool cap_step, cudaEvent_t e0, cudaEvent_t e1) {
  ks_ratio<<<ratio_grid(h), kRatioThreads, 0, h->stream>>>(h->ctls, h->col0, h->col1, h->bcol, h->m, h->row0,
                                                          h->opt.epsilon, h->opt.inf, h->partials, h->peers,
                                                          h->rank, h->world);
  ks_scale_row<<<cdiv(h->ld, kChunk), kChunk, 0, h->stream>>>(h->ctls, h->T, h->ld, h->m, h->n, h->row0, h->row1,
                                                             h->col0, h->col1, h->opt.epsilon, h->opt.inf,
                                                             h->peers, h->rank, h->world, h->plog, h->log_cap,
                                                             h->pos2var);
  if (cap_step) return;
  if (e0) cudaEventRecord(e0, h->stream);
  dim3 grid(cdiv(h->ld, 4ll * 128), cdiv(h->m + 1, 32));
  ks_update<128, 32, 8, 2><<<grid, 128, 0, h->stream>>>(h->ctls, h->T, h->ld, h->m, h->n, h->row0, h->row1,
                                                       h->peers.rowbuf[h->rank], h->col0, h->col1, h->bcol);
  if (e1) cudaEventRecord(e1, h->stream);
}


// ---- persistent loop ------------------------------------------------------------------------
__global__ void k_convert_next(CtlS* ctl, int to_nx) {
  const int slot = (int)((ctl->base.npivots + 1) & 1);
  if (to_nx) ctl->e_nx[slot] = ctl->base.e_next;
  else ctl->base.e_next = ctl->e_nx[slot];
}

int ensure_next_fmt(lps_handle h, bool want_nx) {
  if (!h->next_valid || h->next_in_nx == want_nx) return LPS_OK;
  k_convert_next<<<1, 1, 0, h->stream>>>(h->ctls, want_nx ? 1 : 0);
  CK(cudaGetLastError());
  h->next_in_nx = want_nx;
  return LPS_OK;
}

// tableau bytes per rank, computed from quantities every rank agrees on (the loop shape must be
// the same on all ranks of a sharded solve)
double shard_bytes(lps_handle h) {
  const long long rows = h->sharded ? (h->m_total / h->world) : h->m;
  return 8.0 * (double)(rows + 1) * (double)h->ld;
}

bool use_persistent(lps_handle h) {
  if (h->opt.loop_mode == 1) return false;
  if (h->opt.loop_mode >= 2) return true;
  // auto (profiles/r01_loop_modes.md): the persistent loop wins while the launch chain is a
  // visible share of a pivot; on multi-GB tableaus the hardware CTA scheduler streams ~1.5 %
  // faster than the in-kernel tile queue
  return shard_bytes(h) <= 2.5e9;
}

template <bool kSharded, int kU, int kB>
int launch_loop_t(lps_handle h, const LoopArgs& la_in) {
  if (h->loop_grid == 0) {
    int coop = 0, nb = 0;
    CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->dev));
    if (!coop) return fail(h, LPS_ERR_STATE, "device does not support cooperative launch");
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_loop<kSharded, kU, kB>, kGroupThreads * kB, 0));
    if (nb < 1) return fail(h, LPS_ERR_STATE, "k_loop does not fit on an SM");
    h->loop_grid = h->sm_count;  // one CTA per SM
  }
  LoopArgs la = la_in;
  void* args[] = {&la};
  CK(cudaLaunchCooperativeKernel((void*)k_loop<kSharded, kU, kB>, dim3(h->loop_grid), dim3(kGroupThreads * kB), args,
                                 0, h->stream));
  return LPS_OK;
}

int launch_loop(lps_handle h) {
  LoopArgs la;
  la.ctl = h->ctls;
  la.T = h->T;
  la.ld = h->ld;
  la.mloc = h->m;
  la.n = h->n;
  la.row0 = h->sharded ? h->row0 : 0;
  la.row1 = h->sharded ? h->row1 : h->m;
  la.col0 = h->col0;
  la.col1 = h->col1;
  la.bcol = h->bcol;
  la.rowbuf = h->rowbuf;
  la.partials = h->partials;
  la.plog = h->plog;
  la.log_cap = h->log_cap;
  la.pos2var = h->pos2var;
  la.eps = h->opt.epsilon;
  la.inf = h->opt.inf;
  la.peers = h->peers;
  la.rank = h->rank;
  la.world = h->world;
  const int shape = h->opt.loop_mode;  // 2 = default shape; 3, 4 = tuning alternatives
  if (h->sharded) {
    if (shape == 3) return launch_loop_t<true, 8, 2>(h, la);
    if (shape == 4) return launch_loop_t<true, 4, 4>(h, la);
    return launch_loop_t<true, 8, 3>(h, la);
  }
  if (shape == 3) return launch_loop_t<false, 8, 2>(h, la);
  if (shape == 4) return launch_loop_t<false, 4, 4>(h, la);
  return launch_loop_t<false, 8, 3>(h, la);
}


// ---- blocked loop -----------------------------------------------------------------------------
bool use_blocked(lps_handle h) {
  if (h->block <= 1 || !h->comm || !h->acols) return false;
  if (h->opt.loop_mode != 0) return h->opt.loop_mode >= 5;   // an explicitly requested loop shape wins
  // below ~L2 size the pass is not the cost; the persistent pivot-per-pass loop has the shorter chain
  return shard_bytes(h) > 64e6;
}

bool sweep_available(lps_handle h);
int launch_sweep(lps_handle h);

template <int kLanes, int kU, int kG, bool kPre>
int launch_flush_t(lps_handle h) {
  constexpr int kCH = kU * kLanes * kG;
  const size_t smem = (size_t)h->block * 2 * (kStripCols + kCH) * sizeof(double);
  auto kern = kb_flush<kLanes, kU, kG, kPre>;
  // the shared-memory opt-in belongs to THIS instantiation (and device): a handle reloaded with an LP of another
  // size can land on another instantiation, so the (cheap) attribute call is made at every launch
  {
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int nb = 1;
    if (ce == cudaSuccess && h->flush_grid == 0)
      ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kFlushThreads * kLanes, smem);
    if (ce != cudaSuccess || nb < 1) {
      cudaGetLastError();
      return fail(h, LPS_ERR_STATE, "kb_flush does not fit on an SM with this block_pivots");
    }
    h->flush_grid = h->sm_count;   // persistent: one CTA per SM
  }
  kern<<<h->flush_grid, kFlushThreads * kLanes, smem, h->stream>>>(h->ctls, h->T, h->ld, h->m, h->acols, h->apitch,
                                                                 h->peers.rowbuf[h->rank], h->block);
  return LPS_OK;
}

int launch_flush(lps_handle h) {
  // update_variant >= 10: the TMA pipeline of lps_sweep.cuh, in place; otherwise the cp.async kernel
  // kb_flush<128-thread row-group lanes per CTA, rows per group, groups per lane per chunk, L2 prefetch of the next group>
  if (h->opt.update_variant >= 10 && sweep_available(h)) return launch_sweep(h);
  switch (h->opt.update_variant) {
    default: {
      // chunk height: 256 rows is the best of the B200 sweep (profiles/) while every CTA still gets a
      // few dozen chunks; shorter shards (8-way sharding of the 20,000-row LP) take shorter chunks, and
      // beyond 16 pending pivots 256-row chunks no longer fit in shared memory
      const long long strips = (h->ld + kStripCols - 1) / kStripCols;
      const long long per_cta_256 = strips * ((h->m + 256) / 256) / std::max(1, h->sm_count);
      if (per_cta_256 < 12) return launch_flush_t<4, 4, 4, true>(h);
      if (per_cta_256 < 40 || h->block > 16) return launch_flush_t<4, 4, 8, true>(h);
      return launch_flush_t<4, 4, 16, true>(h);
    }
    case 0: return launch_flush_t<4, 4, 16, true>(h);
    case 1: return launch_flush_t<4, 4, 8, true>(h);
    case 2: return launch_flush_t<4, 4, 8, false>(h);
    case 3: return launch_flush_t<4, 4, 16, false>(h);
    case 4: return launch_flush_t<3, 8, 8, true>(h);
    case 5: return launch_flush_t<3, 8, 8, false>(h);
    case 6: return launch_flush_t<4, 4, 4, true>(h);
    case 7: return launch_flush_t<3, 4, 8, true>(h);
  }
}

// the whole panel of a block (up to `block` pivots) as ONE cooperative launch
bool use_panel_kernel(lps_handle h) {
  if (h->opt.loop_mode == 5) return false;   // 5 = blocked loop with two launches per pivot
  if (h->panel_grid == 0) {
    int coop = 0, nb = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->dev);
    const size_t smem = (size_t)kPanelMax * kPanelThreads * sizeof(double);
    cudaError_t ce = h->sharded
        ? cudaFuncSetAttribute(kb_panel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
        : cudaFuncSetAttribute(kb_panel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce == cudaSuccess)
      ce = h->sharded ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kb_panel<true>, kPanelThreads, smem)
                      : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kb_panel<false>, kPanelThreads, smem);
    if (ce == cudaSuccess && coop && nb >= 1 && h->psync) {
      h->panel_grid = h->sm_count;   // the look-ahead loop already made the slots and sync words
    } else if (ce == cudaSuccess && coop && nb >= 1 &&
        cudaMalloc(&h->ppartials, (size_t)h->sm_count * 128) == cudaSuccess &&
        cudaMalloc(&h->pmins, (size_t)h->sm_count * 128) == cudaSuccess &&
        cudaMemset(h->ppartials, 0, (size_t)h->sm_count * 128) == cudaSuccess &&
        cudaMemset(h->pmins, 0, (size_t)h->sm_count * 128) == cudaSuccess &&
        cudaMalloc(&h->psync, 1024) == cudaSuccess && cudaMemset(h->psync, 0, 1024) == cudaSuccess) {
      h->panel_grid = h->sm_count;   // one CTA per SM; every rank of a sharded solve uses the same grid
    } else {
      cudaGetLastError();
      h->panel_grid = -1;
    }
  }
  return h->panel_grid > 0;
}

int launch_panel(lps_handle h) {
  PanelArgs pa;
  pa.ctl = h->ctls;
  pa.T = h->T;
  pa.ld = h->ld;
  pa.mloc = h->m;
  pa.n = h->n;
  pa.row0 = h->row0;
  pa.row1 = h->row1;
  pa.Acols = h->acols;
  pa.apitch = h->apitch;
  pa.eps = h->opt.epsilon;
  pa.inf = h->opt.inf;
  pa.partials = h->ppartials;
  pa.mins = h->pmins;
  pa.syncw = h->psync;
  pa.gwin = reinterpret_cast<PeerCand*>(h->psync + 192);   // its own 128-byte line behind the five sync words
  pa.ll_off = h->ll_off;
  // 64 tags per launch (two per pivot, at most 32 pivots); tag 0 is the cleared state
  h->panel_launches += 1;
  pa.tag0 = h->panel_launches * 64u;
  pa.peers = h->peers;
  pa.rank = h->rank;
  pa.world = h->world;
  pa.plog = h->plog;
  pa.log_cap = h->log_cap;
  pa.pos2var = h->pos2var;
  pa.block = h->block;
  void* args[] = {&pa};
  const void* fn = h->sharded ? (const void*)kb_panel<true> : (const void*)kb_panel<false>;
  CK(cudaLaunchCooperativeKernel(fn, dim3(h->panel_grid), dim3(kPanelThreads), args,
                                 (size_t)kPanelMax * kPanelThreads * sizeof(double), h->stream));
  return LPS_OK;
}

void launch_panel_step(lps_handle h) {
  const int colgrid = std::max(1, std::min(cdiv(h->m + 1, kColThreads), 4096));
  kb_col<<<colgrid, kColThreads, 0, h->stream>>>(h->ctls, h->T, h->ld, h->m, h->n, h->row0, h->acols, h->apitch,
                                                  h->peers.rowbuf[h->rank], h->opt.epsilon, h->opt.inf,
                                                  h->partials, h->peers, h->rank, h->world);
  kb_row<<<cdiv(h->ld, kChunk), kChunk, 0, h->stream>>>(h->ctls, h->T, h->ld, h->m, h->n, h->row0, h->row1, h->acols,
                                                       h->apitch, h->opt.epsilon, h->opt.inf, h->peers, h->rank,
                                                       h->world, h->plog, h->log_cap, h->pos2var);
}


// ---- TMA pass (lps_sweep.cuh) and look-ahead loop (lps_step.cuh) ------------------------------
// cuTensorMapEncodeTiled comes from the driver through the runtime (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    cudaGetLastError();
  }
  return fn;
}

// row-major FP64 matrix [outer][pitch] seen as a 2-D tensor {inner, outer}; box {box_in, box_out}
bool make_map(CUtensorMap* map, const double* base, unsigned long long inner, unsigned long long outer,
              unsigned long long pitch, unsigned int box_in, unsigned int box_out) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t gdim[2] = {inner, outer};
  const cuuint64_t gstride[1] = {pitch * sizeof(double)};
  const cuuint32_t box[2] = {box_in, box_out};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Shapes of the TMA pass the library is built with (SweepShape, lps_sweep.cuh):
// <pending pivots in registers, rows per thread per sub-pass, columns per thread, consumer warps, warps across the
//  strip, sub-passes per stage>.  pass_shape(): update_variant 10.. pick one explicitly (tuning), otherwise the
// default for the block size.
using ShapeA = SweepShape<16, 2, 2, 12, 4, 2>;   // 512 threads / 128 registers: 12 consumer warps, 256-column strips, 12-row stages
using ShapeB = SweepShape<16, 2, 2, 15, 3, 2>;   // 512 threads / 128 registers: 15 consumer warps, 192-column strips, 20-row stages
using ShapeC = SweepShape<16, 4, 2, 8, 4, 1>;    // 384 threads / 168 registers:  8 consumer warps, 256-column strips,  8-row stages
using ShapeD = SweepShape<16, 4, 2, 8, 4, 2>;    // ... 16-row stages (two sub-passes per barrier round trip)
using ShapeE = SweepShape<16, 4, 2, 8, 4, 1, true>;   // ... software-pipelined over two register sets
using ShapeS = SweepShape<8, 4, 2, 12, 4, 1>;    // small blocks (<= 8 pending pivots)

int pass_shape(lps_handle h) {
  if (h->block <= 8) return 5;
  switch (h->opt.update_variant) {
    case 10: return 0;
    case 11: return 1;
    case 12: return 2;
    case 14: return 4;
    default: return 3;   // ShapeD: the fastest of the B200 sweep (profiles/r02_pass_shapes.md)
  }
}
#define LPS_WITH_SHAPE(h_, ...)                    \
  do {                                             \
    switch (pass_shape(h_)) {                      \
      case 1: { using Shape = ShapeB; __VA_ARGS__; } break;  \
      case 2: { using Shape = ShapeC; __VA_ARGS__; } break;  \
      case 3: { using Shape = ShapeD; __VA_ARGS__; } break;  \
      case 4: { using Shape = ShapeE; __VA_ARGS__; } break;  \
      case 5: { using Shape = ShapeS; __VA_ARGS__; } break;  \
      default: { using Shape = ShapeA; __VA_ARGS__; } break; \
    }                                              \
  } while (0)

int sweep_ks(lps_handle h) { return h->block <= 8 ? 8 : 16; }
bool sweep_available(lps_handle h) {
  return h->block > 1 && h->block <= 16 && h->comm && h->acols && encode_fn() != nullptr;
}
int pass_stage_rows(lps_handle h) { int r = 0; LPS_WITH_SHAPE(h, r = Shape::kSR); return r; }
int pass_cols(lps_handle h) { int r = 0; LPS_WITH_SHAPE(h, r = Shape::kCols); return r; }
int pass_threads(lps_handle h) { int r = 0; LPS_WITH_SHAPE(h, r = Shape::kThreads); return r; }
size_t pass_smem_bytes(lps_handle h) { size_t r = 0; LPS_WITH_SHAPE(h, r = Shape::kBytes); return r; }

size_t step_smem_bytes(lps_handle h) {
  return std::max(pass_smem_bytes(h), (size_t)kLookMax * pass_threads(h) * sizeof(double));
}

// tensor maps of the tableau buffer(s) and of the pending sets; `two` = second tableau buffer present
int ensure_maps(lps_handle h, bool two) {
  const int want = two ? 2 : 1;
  if (h->tm_valid && h->tm_rows_set >= want) return LPS_OK;
  const unsigned int bw = (unsigned int)std::min<long long>(pass_cols(h), h->ld);
  const unsigned int bu = (unsigned int)std::min(sweep_ks(h), h->block);
  const unsigned int sr = (unsigned int)pass_stage_rows(h);
  bool ok = make_map(&h->tm_T[0], h->T, h->ld, h->m + 1, h->ld, bw, sr);
  ok = ok && make_map(&h->tm_T[1], two ? h->T2 : h->T, h->ld, h->m + 1, h->ld, bw, sr);
  for (int s2 = 0; s2 < 2 && ok; s2++) {
    ok = make_map(&h->tm_A[s2], h->acols + (size_t)s2 * h->block * h->apitch, h->apitch, h->block, h->apitch, sr, bu) &&
         make_map(&h->tm_R[s2], h->peers.rowbuf[h->rank] + (size_t)s2 * h->block * h->ld, h->ld, h->block, h->ld, bw, bu);
  }
  if (!ok) return fail(h, LPS_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  h->tm_valid = true;
  h->tm_rows_set = want;
  return LPS_OK;
}

template <typename K>
int step_kernel_setup(lps_handle h, K kern, int threads, size_t smem) {
  cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int nb = 0;
  if (ce == cudaSuccess) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem);
  if (ce != cudaSuccess || nb < 1) {
    cudaGetLastError();
    return fail(h, LPS_ERR_STATE, "the TMA pass kernel does not fit on an SM");
  }
  return LPS_OK;
}

int sweep_chunk_rows(lps_handle h, int ncta) {
  int cr = h->opt.pass_chunk_rows;
  if (cr <= 0) {
    // about two dozen chunks per CTA, so the tail of the pass (CTAs finishing at different times) stays small
    const long long bw = std::min<long long>(pass_cols(h), h->ld);
    const long long nstrips = (h->ld + bw - 1) / bw;
    cr = (int)(((long long)(h->m + 1) * nstrips) / (24ll * std::max(1, ncta)));
    cr = std::min(cr, 120);   // taller chunks let the CTAs drift apart: 120 beat 240 by 5 % on the 20,000-row LP
  }
  const int sr = pass_stage_rows(h);
  cr = std::max(sr, (cr / sr) * sr);
  return cr;
}

void fill_sweep_args(lps_handle h, SweepArgs& sw, int q, bool inplace, int cta0, int ncta) {
  sw.ctl = h->ctls;
  sw.Tbuf[0] = h->T;
  sw.Tbuf[1] = inplace ? h->T : h->T2;
  sw.ld = h->ld;
  sw.rows = h->m + 1;
  sw.chunk_rows = sweep_chunk_rows(h, ncta);
  sw.bw = (int)std::min<long long>(pass_cols(h), h->ld);
  sw.bu = std::min(sweep_ks(h), h->block);
  sw.q = q;
  sw.inplace = inplace ? 1 : 0;
  sw.cta0 = cta0;
  sw.ncta = ncta;
}

// the pass alone, in place, after kb_panel / kb_row (loop modes 5 and 6)
int launch_sweep(lps_handle h) {
  int rc = ensure_maps(h, false);
  if (rc) return rc;
  const size_t smem = pass_smem_bytes(h);
  if (h->sweep_grid == 0) {
    LPS_WITH_SHAPE(h, rc = step_kernel_setup(h, kb_sweep<Shape>, Shape::kThreads, smem));
    if (rc) return rc;
    h->sweep_grid = h->sm_count;
  }
  SweepArgs sw;
  fill_sweep_args(h, sw, 0, true, 0, h->sweep_grid);
  LPS_WITH_SHAPE(h, (kb_sweep<Shape><<<h->sweep_grid, Shape::kThreads, smem, h->stream>>>(sw, h->tm_T[0], h->tm_T[0],
                                                                                        h->tm_A[0], h->tm_R[0])));
  return LPS_OK;
}

// look-ahead step with the cp.async pass (flush_role): chunk height as launch_flush picks it
int step_flush_kg(lps_handle h, int ncta) {
  const long long strips = (h->ld + kStripCols - 1) / kStripCols;
  const long long per_cta_256 = strips * ((h->m + 256) / 256) / std::max(1, ncta);
  if (per_cta_256 < 12) return 4;
  if (per_cta_256 < 40) return 8;
  return 16;
}
// the look-ahead loop's pass role (update_variant 0..9: the cp.async kernel, >= 10: a shape of the TMA pipeline)
bool step_uses_flush(lps_handle h) {
  if (h->opt.update_variant >= 10) return false;
  if (h->opt.update_variant >= 0) return true;
  // default: the TMA pipeline (shape D) on multi-GB shards — there the GPU runs against its power cap for seconds
  // and the TMA pass, with half the shared-memory operand traffic, keeps the higher clock (5.6 - 6.0 k pivots/s
  // against 5.3 - 5.4 k on the 20,000 x 40,000 LP, sustained); the cp.async pass on smaller shards, where it is
  // 10 - 15 % faster per step (profiles/r02_summary.md)
  return shard_bytes(h) < 2.5e9;
}
size_t step_flush_smem(lps_handle h, int kg) {
  const size_t pass = (size_t)h->block * 2 * (kStripCols + 4 * 4 * kg) * sizeof(double);
  return std::max(pass, (size_t)kLookMax * 512 * sizeof(double));
}

// loop_mode 8: the warp-specialised look-ahead step (kb_step_ws): pass and panel warps in every CTA
bool step_is_ws(lps_handle h) { return h->opt.loop_mode == 8; }
int step_ws_kg(lps_handle h) {
  const long long strips = (h->ld + kStripCols - 1) / kStripCols;
  const long long per_cta_192 = strips * ((h->m + 192) / 192) / std::max(1, h->sm_count);
  if (per_cta_192 < 12) return 4;
  if (per_cta_192 < 40) return 8;
  return 16;
}
size_t step_ws_pass_bytes(lps_handle h, int kg) { return (size_t)h->block * 2 * (kStripCols + 12 * kg) * sizeof(double); }
size_t step_ws_smem(lps_handle h, int kg) {
  (void)h; (void)kg;
  return (size_t)220 * 1024;      // the pass's operand slices, then the panel's staging area takes the rest
}

bool use_look(lps_handle h) {
  if (!sweep_available(h)) return false;
  if (h->opt.loop_mode == 7 || h->opt.loop_mode == 8) return true;
  return h->opt.loop_mode == 0 && shard_bytes(h) > 64e6;
}

// The a-priori split of the look-ahead step (lps_plan_split_model in the C ABI; no device involved).  Both roles
// are throughput-bound on the SMs they get (measured on B200, profiles/r02_summary.md):
//   pass   ~0.37 ms per GB of shard on the whole GPU (16 pivots replayed) + 0.1 ms of tail on small shards,
//          proportionally slower on fewer SMs;
//   panel  ~10 us of syncs per pivot + 11.5 us per 1000 cells (local rows + columns) that one of its CTAs has
//          to replay; sharded (two NVLink hops): ~19 us + 10.3 us per 1000 cells (14.9 with one cell per thread, 8 ranks).
// Pick the split that minimises the slower of the two.  It only seeds the first run of a handle (tune_split).
int split_model(int grid, int block, int world, long long rows_local, long long ld) {
  const double shard = 8.0 * (double)(rows_local + 1) * (double)ld;
  const double pass_ms_full = (0.367 * shard / 1e9 + 0.11) * block / 16.0;
  double best = 1e30;
  int P = std::min(8, std::max(1, grid - 1));
  for (int p = 2; p <= grid / 2; p++) {
    const double kcells = 1e-3 * ((double)(rows_local + 1) + (double)ld) / p;
    const double panel_ms = block * (world >= 8 ? 19.0 + 14.9 * kcells : world > 1 ? 19.0 + 10.3 * kcells
                                                                                    : 10.0 + 11.5 * kcells) * 1e-3;
    const double pass_ms = pass_ms_full * grid / (double)(grid - p);
    const double step = std::max(panel_ms, pass_ms);
    if (step < best) { best = step; P = p; }
  }
  return std::max(1, std::min(P, grid - 1));
}

// The re-fit after a run (lps_plan_split_tuned in the C ABI): panel(p) = block * (A + B / p) with the sync share A
// fixed and B from the measured panel clock at `cur` CTAs; pass(p) = the measured pass scaled by the CTAs it had.
// Moves at most a third of the way per run and only for a predicted gain above 2 %.
int split_tuned(int grid, int block, int world, int cur, double panel_us_per_pivot, double pass_us_per_block) {
  if (cur < 1 || cur >= grid || !(panel_us_per_pivot > 0.0) || !(pass_us_per_block > 0.0)) return cur;
  const double A = std::min(world > 1 ? 19.0 : 10.0, 0.6 * panel_us_per_pivot);
  const double B = (panel_us_per_pivot - A) * cur;
  const double full = pass_us_per_block * (grid - cur) / grid;
  auto cost = [&](int p) { return std::max(block * (A + B / p), full * grid / (double)(grid - p)); };
  const int reach = std::max(2, cur / 3);
  int best = cur;
  for (int p = std::max(2, cur - reach); p <= std::min(grid / 2, cur + reach); p++)
    if (cost(p) < cost(best)) best = p;
  return cost(best) < 0.98 * cost(cur) ? best : cur;
}

int look_panel_ctas(lps_handle h) {
  int P = h->opt.panel_ctas;
  if (P <= 0 && h->tuned_P > 0 && h->tuned_m == h->m && h->tuned_ld == h->ld) P = h->tuned_P;     // tune_split()
  if (P <= 0) P = split_model(h->sm_count, h->block, h->world, h->sharded ? h->m_total / h->world : h->m, h->ld);
  return std::max(1, std::min(P, h->sm_count - 1));
}

// After a run: re-fit the panel / pass split from the two roles' own clocks (CtlS::dbg_ns, written by the panel's
// scribe and by the last pass CTA to retire).  The model only has to be right enough for the first run of a
// handle; every later run starts from what the previous one measured on this GPU, at this shard size, with these
// peers.  The pivots do not depend on the split (tests: panel_ctas sweep).
void tune_split(lps_handle h) {
  if (h->opt.panel_ctas > 0 || step_is_ws(h)) return;
  if (const char* tv = std::getenv("LPS_SPLIT_TUNE")) if (tv[0] == '0') return;
  const unsigned long long* d = h->h_ctls->dbg_ns;
  if (d[15] < (unsigned long long)(2 * h->block) || d[11] < 2) return;
  const int P0 = look_panel_ctas(h);
  h->tuned_P = split_tuned(h->step_grid, h->block, h->world, P0, (double)d[14] / (double)d[15] * 1e-3,
                           (double)d[10] / (double)d[11] * 1e-3);
  h->tuned_m = h->m;
  h->tuned_ld = h->ld;
}

// second tableau buffer, running vectors, sync words, tensor maps, kernel attributes
int ensure_look(lps_handle h) {
  const size_t need = (size_t)(h->m + 1) * (size_t)h->ld * sizeof(double);
  if (need > h->T2_bytes) {
    if (h->T2) cudaFree(h->T2);
    h->T2 = nullptr;
    h->T2_bytes = 0;
    h->tm_valid = false;
    cudaError_t ce = cudaMalloc(&h->T2, need);
    if (ce != cudaSuccess) {
      cudaGetLastError();
      return fail(h, LPS_ERR_NOMEM, "cudaMalloc(second tableau buffer of the look-ahead loop)", ce);
    }
    h->T2_bytes = need;
  }
  if ((size_t)h->m + 1 > h->bvec_cap) {
    if (h->bvec) cudaFree(h->bvec);
    CK(cudaMalloc(&h->bvec, ((size_t)h->m + 1 + 64) * sizeof(double)));
    h->bvec_cap = (size_t)h->m + 1 + 64;
  }
  if ((size_t)h->ld > h->cvec_cap) {
    if (h->cvec) cudaFree(h->cvec);
    CK(cudaMalloc(&h->cvec, ((size_t)h->ld + 64) * sizeof(double)));
    h->cvec_cap = (size_t)h->ld + 64;
  }
  if (!h->psync) {
    CK(cudaMalloc(&h->ppartials, (size_t)h->sm_count * 128));
    CK(cudaMalloc(&h->pmins, (size_t)h->sm_count * 128));
    CK(cudaMalloc(&h->psync, 1024));
    CK(cudaMemsetAsync(h->ppartials, 0, (size_t)h->sm_count * 128, h->stream));
    CK(cudaMemsetAsync(h->pmins, 0, (size_t)h->sm_count * 128, h->stream));
    CK(cudaMemsetAsync(h->psync, 0, 1024, h->stream));
  }
  int rc = ensure_maps(h, true);
  if (rc) return rc;
  if (h->step_grid == 0 && (step_uses_flush(h) || step_is_ws(h))) {
    int coop = 0;
    CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->dev));
    if (!coop) return fail(h, LPS_ERR_STATE, "device does not support cooperative launch");
    h->step_grid = h->sm_count;     // the attribute is set per launch (the instantiation depends on the shape)
  }
  if (h->step_grid == 0) {
    int coop = 0;
    CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->dev));
    if (!coop) return fail(h, LPS_ERR_STATE, "device does not support cooperative launch");
    const size_t smem = step_smem_bytes(h);
    if (h->sharded) LPS_WITH_SHAPE(h, rc = step_kernel_setup(h, kb_step<true, Shape>, Shape::kThreads, smem));
    else LPS_WITH_SHAPE(h, rc = step_kernel_setup(h, kb_step<false, Shape>, Shape::kThreads, smem));
    if (rc) return rc;
    h->step_grid = h->sm_count;
  }
  return LPS_OK;
}

int launch_step(lps_handle h) {
  StepArgs sa;
  const int q = (int)(h->look_launches & 1u);
  // the first launch of a run has no pending block to apply: all CTAs take the panel role
  const bool ws = step_is_ws(h);
  const int P = (ws || h->look_launches == 0) ? h->step_grid : look_panel_ctas(h);
  fill_sweep_args(h, sa.sw, q, false, ws ? 0 : P, ws ? h->step_grid : h->step_grid - P);
  sa.mloc = h->m;
  sa.n = h->n;
  sa.row0 = h->row0;
  sa.row1 = h->row1;
  sa.panel_ctas = P;
  sa.block = h->block;
  sa.Acols = h->acols;
  sa.apitch = h->apitch;
  sa.bvec = h->bvec;
  sa.cvec = h->cvec;
  sa.eps = h->opt.epsilon;
  sa.inf = h->opt.inf;
  sa.partials = h->ppartials;
  sa.mins = h->pmins;
  sa.syncw = h->psync;
  sa.gwin = reinterpret_cast<PeerCand*>(h->psync + 192);
  sa.ll_off = h->ll_off;
  sa.peers = h->peers;
  sa.rank = h->rank;
  sa.world = h->world;
  sa.plog = h->plog;
  sa.log_cap = h->log_cap;
  sa.pos2var = h->pos2var;
  h->panel_launches += 1;
  sa.tag0 = h->panel_launches * 64u;
  h->look_launches += 1;
  {
    // two rows / columns per panel thread and trip (16-byte staging slots: fewer trips — 135 -> 87 us per pivot at 8
    // CTAs on one GPU, 58 -> 49 us on 4 ranks); one (8-byte slots) on 8 ranks, where the shorter trips let the
    // owner's first packets leave earlier (43 -> 40 us).  Forms that did not pay (profiles/r02_summary.md): trips
    // pipelined without CTA barriers, operands staged by cp.async.bulk, pending rows pinned in L2.
    const char* cv = std::getenv("LPS_PANEL_CELLS");
    sa.cells = cv ? std::max(1, std::min(2, std::atoi(cv))) : (h->world >= 8 ? 1 : 2);
  }
  {
    const char* hv = std::getenv("LPS_L2_HINTS");
    sa.hints = hv ? std::atoi(hv) : 0;
  }
  sa.stage_doubles = (int)((step_uses_flush(h) ? step_flush_smem(h, step_flush_kg(h, h->step_grid - P)) : step_smem_bytes(h)) / sizeof(double));
  if (ws) {
    const int kg = step_ws_kg(h);
    const size_t smem = step_ws_smem(h, kg);
    sa.stage_doubles = (int)((smem - step_ws_pass_bytes(h, kg)) / sizeof(double));
    const void* wfn;
    if (h->sharded) wfn = kg == 16 ? (const void*)kb_step_ws<true, 16> : kg == 8 ? (const void*)kb_step_ws<true, 8> : (const void*)kb_step_ws<true, 4>;
    else wfn = kg == 16 ? (const void*)kb_step_ws<false, 16> : kg == 8 ? (const void*)kb_step_ws<false, 8> : (const void*)kb_step_ws<false, 4>;
    CK(cudaFuncSetAttribute(wfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* wargs[] = {&sa};
    CK(cudaLaunchCooperativeKernel(wfn, dim3(h->step_grid), dim3(512), wargs, smem, h->stream));
    return LPS_OK;
  }
  if (step_uses_flush(h)) {
    const int kg = step_flush_kg(h, h->step_grid - P);
    const size_t smem = step_flush_smem(h, kg);
    const void* ffn;
    if (h->sharded) ffn = kg == 16 ? (const void*)kb_step_flush<true, 4, 4, 16, true> : kg == 8 ? (const void*)kb_step_flush<true, 4, 4, 8, true> : (const void*)kb_step_flush<true, 4, 4, 4, true>;
    else ffn = kg == 16 ? (const void*)kb_step_flush<false, 4, 4, 16, true> : kg == 8 ? (const void*)kb_step_flush<false, 4, 4, 8, true> : (const void*)kb_step_flush<false, 4, 4, 4, true>;
    CK(cudaFuncSetAttribute(ffn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* fargs[] = {&sa};
    CK(cudaLaunchCooperativeKernel(ffn, dim3(h->step_grid), dim3(4 * kFlushThreads), fargs, smem, h->stream));
    return LPS_OK;
  }
  void* args[] = {&sa, &h->tm_T[0], &h->tm_T[1], &h->tm_A[q], &h->tm_R[q]};
  const void* fn = nullptr;
  if (h->sharded) LPS_WITH_SHAPE(h, fn = (const void*)kb_step<true, Shape>);
  else LPS_WITH_SHAPE(h, fn = (const void*)kb_step<false, Shape>);
  CK(cudaLaunchCooperativeKernel(fn, dim3(h->step_grid), dim3(pass_threads(h)), args, step_smem_bytes(h), h->stream));
  return LPS_OK;
}

int run_look(lps_handle h, int64_t max_pivots, lps_run_result* res) {
  const long long start_pivots = h->total_pivots;
  const int S = h->block;
  long long launches = 0;
  if (h->panel_dirty && h->psync) {
    CK(cudaMemsetAsync(h->psync, 0, 1024, h->stream));
    h->panel_dirty = false;
  }
  CK(cudaEventRecord(h->ev_begin, h->stream));
  // the tableau is fully applied between calls: the entering column and the running b column /
  // objective row come straight from it
  ks_begin_run<<<1, 1, 0, h->stream>>>(h->ctls, max_pivots, 1);
  ks_first_positive<<<std::max(1, cdiv(h->n, 256)), 256, 0, h->stream>>>(h->ctls, h->T + (long long)h->m * h->ld, h->n,
                                                                        h->opt.epsilon);
  kb_init_vec<<<std::max(1, cdiv(std::max<long long>(h->m + 1, h->ld), 256)), 256, 0, h->stream>>>(
      h->T, h->ld, h->m, h->n, h->bvec, h->cvec);
  launches += 3;
  h->look_launches = 0;
  // blocks per host check: about 30 ms of device work
  const double bytes = 16.0 * (double)(h->m + 1) * (double)(h->n + 1);
  const double est_us = bytes / 5.0e6 + 16.0 * S;
  long long batch = std::max(2ll, std::min((long long)(30000.0 / est_us), 256ll));
  const bool timed = h->opt.time_kernels != 0;
  if (timed) {
    while ((long long)h->ev.size() < 2 * (batch + 1)) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      h->ev.push_back(e);
    }
  }
  double upd_ms = 0.0;
  long long upd_launches = 0, run_index = 0;     // run_index: launches of this run so far
  long long remaining = (max_pivots < 0) ? -1 : (long long)max_pivots;
  int rc = LPS_OK;
  for (;;) {
    // a capped run needs ceil(remaining / S) + 1 launches: the panel runs one block ahead of the pass
    long long todo = batch;
    if (remaining >= 0) todo = std::min(batch, (remaining + S - 1) / S + 1);
    if (h->h_ctl->status != kRunning && run_index > 0) todo = 1;        // draining the last pending block
    for (long long k = 0; k < todo; k++) {
      if (timed) cudaEventRecord(h->ev[2 * k], h->stream);
      rc = launch_step(h);
      if (rc) return rc;
      if (timed) cudaEventRecord(h->ev[2 * k + 1], h->stream);
      launches++;
    }
    CK(cudaGetLastError());
    rc = sync_ctl(h);
    if (rc) return rc;
    const long long done_now = h->h_ctl->npivots - h->total_pivots;
    // passes happen in launches 1 .. sweeps_done of the run (launch 0 has nothing to apply yet)
    const long long swept = (long long)h->h_ctls->sweeps_done;
    for (long long k = 0; k < todo; k++) {
      const long long r = run_index + k;
      if (r >= 1 && r <= swept) {
        if (timed) {
          float ms = 0.f;
          if (cudaEventElapsedTime(&ms, h->ev[2 * k], h->ev[2 * k + 1]) == cudaSuccess) upd_ms += ms;
        }
        upd_launches++;
      }
    }
    run_index += todo;
    h->total_pivots = h->h_ctl->npivots;
    if (remaining >= 0) remaining -= done_now;
    if (h->h_ctl->status != kRunning && (h->h_ctls->blk_pend[0] | h->h_ctls->blk_pend[1]) == 0) break;
  }
  if (const char* dbg = std::getenv("LPS_DEBUG")) {
    if (dbg[0] == '1') {
      const unsigned long long* d = h->h_ctls->dbg_ns;
      const double per = d[15] ? 1e-3 / (double)d[15] : 0.0;
      std::fprintf(stderr, "lps look-ahead run: rank %d  panel role: %.1f us per pivot over %llu pivots (%d CTAs); "
                           "%lld launches, %lld with a pass;  per pivot: column trips %.1f | sync A %.1f | gather+exchange %.1f | "
                           "row trips %.1f | sync B %.1f | gather+commit %.1f us;  pass role: %.1f us per block;  "
                           "first launch (all CTAs on the panel): %.1f us per pivot\n",
                   h->rank, (double)d[14] * per, d[15], look_panel_ctas(h), launches, upd_launches, (double)d[0] * per,
                   (double)d[1] * per, (double)d[2] * per, (double)d[3] * per, (double)d[4] * per, (double)d[5] * per,
                   d[11] ? (double)d[10] * 1e-3 / (double)d[11] : 0.0, d[13] ? (double)d[12] * 1e-3 / (double)d[13] : 0.0);
    }
  }
  tune_split(h);
  // the tableau may have ended up in the second buffer: make it the handle's current one
  if (h->h_ctls->cur_at[h->look_launches & 1u] == 1) {
    std::swap(h->T, h->T2);
    std::swap(h->T_bytes, h->T2_bytes);
    std::swap(h->tm_T[0], h->tm_T[1]);
  }
  if (h->h_ctl->status == kCommTimeout) {
    h->panel_dirty = true;
    return fail(h, LPS_ERR_COMM, "shard: timed out waiting for a peer rank");
  }
  CK(cudaEventRecord(h->ev_end, h->stream));
  CK(cudaEventSynchronize(h->ev_end));
  h->next_valid = false;
  h->col_holds = -1;
  if (res) {
    std::memset(res, 0, sizeof(*res));
    res->verdict = h->h_ctl->status;
    res->last_entering = h->h_ctl->e_cur;
    res->last_leaving = h->h_ctl->l_cur;
    res->npivots = h->total_pivots - start_pivots;
    res->total_pivots = h->total_pivots;
    double corner = 0.0;
    CK(cudaMemcpy(&corner, h->T + (long long)h->m * h->ld + h->n, sizeof(double), cudaMemcpyDeviceToHost));
    res->v = 0.0 - corner;
    cudaEventElapsedTime(&res->device_ms, h->ev_begin, h->ev_end);
    res->update_ms = (float)upd_ms;
    res->update_launches = upd_launches;
    res->kernel_launches = launches;
  }
  return LPS_OK;
}

int run_blocked(lps_handle h, int64_t max_pivots, lps_run_result* res) {
  const long long start_pivots = h->total_pivots;
  const int S = h->block;
  const bool panel_kernel = use_panel_kernel(h);
  long long launches = 0;
  if (h->panel_dirty && h->psync) {
    CK(cudaMemsetAsync(h->psync, 0, 1024, h->stream));
    h->panel_dirty = false;
  }
  CK(cudaEventRecord(h->ev_begin, h->stream));
  // the tableau is fully applied between calls, so the entering column comes from its objective row
  ks_begin_run<<<1, 1, 0, h->stream>>>(h->ctls, max_pivots, 1);
  ks_first_positive<<<std::max(1, cdiv(h->n, 256)), 256, 0, h->stream>>>(h->ctls, h->T + (long long)h->m * h->ld, h->n,
                                                                        h->opt.epsilon);
  launches += 2;
  // pivots per host check: about 30 ms of device work, whole blocks
  const double bytes = 16.0 * (double)(h->m + 1) * (double)(h->n + 1);
  const double est_us = bytes / 5.0e6 / S + 2.0 * (double)(h->m + 1) * (double)(h->n + 1) / 12.0e6 + 14.0;
  long long batch = (long long)(30000.0 / est_us);
  batch = std::max((long long)S, std::min(batch, 4096ll));
  batch = (batch / S) * S;
  const bool timed = h->opt.time_kernels != 0;
  const long long flushes_per_batch = batch / S + 1;
  if (timed) {
    while ((long long)h->ev.size() < 2 * flushes_per_batch) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      h->ev.push_back(e);
    }
  }
  double upd_ms = 0.0;
  long long upd_launches = 0;
  long long remaining = (max_pivots < 0) ? -1 : (long long)max_pivots;
  int rc = LPS_OK;
  for (;;) {
    // one extra panel step past the cap is what turns "cap reached" into a verdict
    const long long todo = (remaining < 0) ? batch : std::min(batch, remaining + 1);
    long long nflush = 0;
    for (long long k = 0; k < todo; k++) {
      if (!panel_kernel) {
        launch_panel_step(h);
        launches += 2;
      } else if (k % S == 0) {
        rc = launch_panel(h);     // this block's pivots (it stops by itself at the cap or a verdict)
        if (rc) return rc;
        launches += 1;
      }
      if ((k + 1) % S == 0 || k == todo - 1) {
        if (timed) cudaEventRecord(h->ev[2 * nflush], h->stream);
        rc = launch_flush(h);
        if (rc) return rc;
        if (timed) cudaEventRecord(h->ev[2 * nflush + 1], h->stream);
        nflush++;
        launches++;
      }
    }
    CK(cudaGetLastError());
    rc = sync_ctl(h);
    if (rc) return rc;
    const long long done_now = h->h_ctl->npivots - h->total_pivots;
    const long long worked = (done_now + S - 1) / S;   // flushes of this batch that had pending pivots
    if (timed) {
      for (long long f = 0; f < worked && f < nflush; f++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev[2 * f], h->ev[2 * f + 1]) == cudaSuccess) upd_ms += ms;
      }
    }
    upd_launches += std::min(worked, nflush);
    h->total_pivots = h->h_ctl->npivots;
    if (remaining >= 0) remaining -= done_now;
    if (h->h_ctl->status != kRunning) break;
  }
  if (h->h_ctl->status == kCommTimeout) {
    h->panel_dirty = true;
    return fail(h, LPS_ERR_COMM, "shard: timed out waiting for a peer rank");
  }
  if ((h->h_ctls->blk_pend[0] | h->h_ctls->blk_pend[1]) != 0) return fail(h, LPS_ERR_STATE, "blocked loop: pivots left pending after the last pass");
  CK(cudaEventRecord(h->ev_end, h->stream));
  CK(cudaEventSynchronize(h->ev_end));
#ifdef LPS_PANEL_TIMING
  {
    const unsigned long long* d = h->h_ctls->dbg_ns;
    std::fprintf(stderr,
                 "kb_panel CTA0 ns: A[loads+sync %llu | replay+div %llu | warp-reduce+sync %llu | publish %llu] gatherA %llu | "
                 "winner %llu | B[loads+scalars+sync %llu | columns %llu | block-min+sync %llu | publish %llu] gatherB %llu | "
                 "commit %llu  (pivots %lld)  SM clock during the panel kernels: %.0f MHz\n",
                 d[6], d[7], d[8], d[0], d[1], d[2], d[9], d[10], d[11], d[3], d[4], d[5], (long long)h->total_pivots,
                 d[13] ? 1e3 * (double)d[12] / (double)d[13] : 0.0);
  }
#endif
  // the pivot-per-pass staging vectors (colbuf, bcol) are not maintained by this loop
  h->next_valid = false;
  h->col_holds = -1;
  if (res) {
    std::memset(res, 0, sizeof(*res));
    res->verdict = h->h_ctl->status;
    res->last_entering = h->h_ctl->e_cur;
    res->last_leaving = h->h_ctl->l_cur;
    res->npivots = h->total_pivots - start_pivots;
    res->total_pivots = h->total_pivots;
    double corner = 0.0;
    CK(cudaMemcpy(&corner, h->T + (long long)h->m * h->ld + h->n, sizeof(double), cudaMemcpyDeviceToHost));
    res->v = 0.0 - corner;
    cudaEventElapsedTime(&res->device_ms, h->ev_begin, h->ev_end);
    res->update_ms = (float)upd_ms;
    res->update_launches = upd_launches;
    res->kernel_launches = launches;
  }
  return LPS_OK;
}

// FP64 issue-rate probe (the roof of the blocked pass beside HBM): the replay's instruction mix on independent
// chains; the multiplier is another chain's running value, so nothing can be hoisted out of the loop
template <int kChains>
__global__ void __launch_bounds__(256, 4) k_fp64_probe(double* out, double r, int iters) {
  double x[kChains];
#pragma unroll
  for (int c = 0; c < kChains; c++) x[c] = 1.0 + 1e-3 * (double)((threadIdx.x + c) & 63);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < kChains; c++) x[c] = __dsub_rn(x[c], __dmul_rn(x[(c + 1) % kChains], r));
  }
  double s2 = 0;
#pragma unroll
  for (int c = 0; c < kChains; c++) s2 += x[c];
  if (s2 == 12345.678) out[0] = s2;      // keeps the chains alive without a store per thread
}

}  // namespace

extern "C" {

int lps_abi_version(void) { return LPS_ABI_VERSION; }

void lps_default_options(lps_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->epsilon = 1e-9;
  o->inf = 1e50;
  o->device = -1;
  o->update_variant = -1;
}

const char* lps_status_string(int s) {
  switch (s) {
    case LPS_OK: return "ok";
    case LPS_ERR_INVALID: return "invalid argument";
    case LPS_ERR_CUDA: return "CUDA error";
    case LPS_ERR_STATE: return "call in wrong state";
    case LPS_ERR_NOMEM: return "out of device memory";
    case LPS_ERR_NODEVICE: return "no CUDA device (there is no CPU fallback)";
    case LPS_ERR_COMM: return "multi-GPU exchange failure";
    default: return "unknown";
  }
}

int lps_create(lps_handle* out, const lps_options* opts) {
  if (!out) return LPS_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return LPS_ERR_NODEVICE;
  lps_handle h = new (std::nothrow) lps_handle_s();
  if (!h) return LPS_ERR_NOMEM;
  if (opts) h->opt = *opts; else lps_default_options(&h->opt);
  // pivots deferred per tableau pass: 0 = default, 1 = off (pivot-per-pass kernels only)
  h->block = (h->opt.block_pivots == 0) ? kDefaultBlock : std::max(1, std::min(h->opt.block_pivots, kFlushMaxBlock));
  if (h->opt.device >= 0) {
    if (h->opt.device >= ndev) { delete h; return LPS_ERR_INVALID; }
    h->dev = h->opt.device;
  } else {
    cudaGetDevice(&h->dev);
  }
  cudaError_t ce = cudaSetDevice(h->dev);
  if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->dev);
  if (ce == cudaSuccess) {
    if (h->opt.stream) {
      h->stream = (cudaStream_t)h->opt.stream;
    } else {
      ce = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
      h->own_stream = true;
    }
  }
  if (ce == cudaSuccess) ce = cudaMalloc(&h->ctls, sizeof(CtlS));
  if (ce == cudaSuccess) ce = cudaMallocHost(&h->h_ctls, sizeof(CtlS));
  if (ce == cudaSuccess) {
    h->ctl = &h->ctls->base;
    h->h_ctl = &h->h_ctls->base;
  }
  if (ce == cudaSuccess) ce = cudaMalloc(&h->partials, 4096 * sizeof(Cand));
  if (ce == cudaSuccess) ce = cudaMalloc(&h->plog, (size_t)h->log_cap * sizeof(int2));
  if (ce == cudaSuccess) ce = cudaEventCreate(&h->ev_begin);
  if (ce == cudaSuccess) ce = cudaEventCreate(&h->ev_end);
  if (ce != cudaSuccess) {
    lps_destroy(h);
    return LPS_ERR_CUDA;
  }
  *out = h;
  return LPS_OK;
}

int lps_destroy(lps_handle h) {
  if (!h) return LPS_OK;
  cudaSetDevice(h->dev);
  if (h->stream) cudaStreamSynchronize(h->stream);
  free_tableau(h);
  for (double* p : {h->col0, h->col1, h->bcol, h->scratch, h->rowbuf}) if (p) cudaFree(p);
  if (h->pos2var) cudaFree(h->pos2var);
  if (h->plog) cudaFree(h->plog);
  for (void* p : h->ipc_opened) cudaIpcCloseMemHandle(p);
  if (h->comm) cudaFree(h->comm);
  if (h->ctls) cudaFree(h->ctls);
  if (h->h_ctls) cudaFreeHost(h->h_ctls);
  if (h->partials) cudaFree(h->partials);
  if (h->d_ops) cudaFree(h->d_ops);
  if (h->acols) cudaFree(h->acols);
  if (h->T2) cudaFree(h->T2);
  if (h->bvec) cudaFree(h->bvec);
  if (h->cvec) cudaFree(h->cvec);
  if (h->ppartials) cudaFree(h->ppartials);
  if (h->pmins) cudaFree(h->pmins);
  if (h->psync) cudaFree(h->psync);
  for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
  if (h->ev_begin) cudaEventDestroy(h->ev_begin);
  if (h->ev_end) cudaEventDestroy(h->ev_end);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return LPS_OK;
}

const char* lps_last_error(lps_handle h) { return h ? h->err.c_str() : "null handle"; }

static int load_common(lps_handle h, int m, int n_src, int n_cols, const double* A, int64_t lda,
                       const double* b, const double* c, double v, bool aux) {
  if (!h) return LPS_ERR_INVALID;
  if (m < 0 || n_src < 0 || (m > 0 && n_src > 0 && !A) || (m > 0 && !b) || lda < n_src)
    return fail(h, LPS_ERR_INVALID, "lps_load: bad dimensions or null buffer");
  CK(cudaSetDevice(h->dev));
  int rc = ensure_buffers(h, m, n_cols);
  if (rc) return rc;
  rc = ensure_self_comm(h);
  if (rc) return rc;
  const long long ld = h->ld;
  CK(cudaMemsetAsync(h->T, 0, (size_t)(m + 1) * ld * sizeof(double), h->stream));
  if (m > 0 && n_src > 0)
    CK(cudaMemcpy2DAsync(h->T, ld * sizeof(double), A, (size_t)lda * sizeof(double),
                         (size_t)n_src * sizeof(double), m, cudaMemcpyHostToDevice, h->stream));
  if (m > 0) {
    CK(cudaMemcpyAsync(h->scratch, b, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    k_set_column<<<cdiv(m, 256), 256, 0, h->stream>>>(h->T, ld, m, n_cols, h->scratch, 0.0, 0);
  }
  if (aux) {
    if (m > 0) k_set_column<<<cdiv(m, 256), 256, 0, h->stream>>>(h->T, ld, m, n_src, nullptr, -1.0, 1);
    const double minus1 = -1.0;  // c_aux = (0,…,0,-1)
    CK(cudaMemcpyAsync(h->T + (long long)m * ld + n_src, &minus1, sizeof(double),
                       cudaMemcpyHostToDevice, h->stream));
  } else {
    if (n_src > 0)
      CK(cudaMemcpyAsync(h->T + (long long)m * ld, c, (size_t)n_src * sizeof(double),
                         cudaMemcpyHostToDevice, h->stream));
    const double negv = -v;
    CK(cudaMemcpyAsync(h->T + (long long)m * ld + n_cols, &negv, sizeof(double),
                       cudaMemcpyHostToDevice, h->stream));
  }
  CK(cudaGetLastError());
  rc = reset_state(h);
  if (rc) return rc;
  CK(cudaStreamSynchronize(h->stream));
  return LPS_OK;
}

int lps_load(lps_handle h, int m, int n, const double* A, int64_t lda, const double* b,
             const double* c, double v) {
  if (h && n > 0 && !c) return fail(h, LPS_ERR_INVALID, "lps_load: null c");
  return load_common(h, m, n, n, A, lda, b, c, v, false);
}

int lps_load_aux(lps_handle h, int m, int n, const double* A, int64_t lda, const double* b) {
  return load_common(h, m, n, n + 1, A, lda, b, nullptr, 0.0, true);
}

int lps_generate_lp(lps_handle h, int kind, int m, int n, uint64_t seed, int param) {
  if (!h || m <= 0 || n <= 0) return fail(h, LPS_ERR_INVALID, "lps_generate_lp: bad dimensions");
  if (kind < 0 || kind > 2 || (kind == LPS_GEN_UNBOUNDED && (param < 0 || param >= n)) ||
      (kind == LPS_GEN_ASSIGNMENT && m < 2))
    return fail(h, LPS_ERR_INVALID, "lps_generate_lp: bad kind / parameter");
  CK(cudaSetDevice(h->dev));
  int rc = ensure_buffers(h, m, n);
  if (rc) return rc;
  rc = ensure_self_comm(h);
  if (rc) return rc;
  dim3 grid(std::min(cdiv(h->ld, 256), 64), std::min(m + 1, 65535));
  ks_generate_lp<<<grid, 256, 0, h->stream>>>(h->T, h->ld, m, n, 0, m, seed, kind, param);
  CK(cudaGetLastError());
  rc = reset_state(h);
  if (rc) return rc;
  CK(cudaStreamSynchronize(h->stream));
  return LPS_OK;
}

int lps_generate_dense(lps_handle h, int m, int n, uint64_t seed, int pos_permille) {
  return lps_generate_lp(h, LPS_GEN_DENSE, m, n, seed, pos_permille);
}

int lps_get_entering(lps_handle h, int* e) {
  if (!h || !e) return LPS_ERR_INVALID;
  if (h->sharded) return fail(h, LPS_ERR_STATE, "not available on a row shard (use lps_run)");
  if (!h->loaded) return fail(h, LPS_ERR_STATE, "nothing loaded");
  CK(cudaSetDevice(h->dev));
  int rc = h->next_valid ? ensure_next_fmt(h, false) : prepare_next(h);
  if (rc) return rc;
  rc = sync_ctl(h);
  if (rc) return rc;
  *e = (h->h_ctl->e_next == kNone) ? -1 : h->h_ctl->e_next;
  return LPS_OK;
}

int lps_get_leaving(lps_handle h, int e, int* l) {
  if (!h || !l) return LPS_ERR_INVALID;
  if (h->sharded) return fail(h, LPS_ERR_STATE, "not available on a row shard (use lps_run)");
  if (!h->loaded) return fail(h, LPS_ERR_STATE, "nothing loaded");
  if (e < 0 || e >= h->n) return fail(h, LPS_ERR_INVALID, "getLeaving: entering out of range");
  CK(cudaSetDevice(h->dev));
  // stage column e and b; this invalidates the loop's staged (e_next, column) pair unless e is it
  k_extract<<<cdiv(h->m + 1, 256), 256, 0, h->stream>>>(h->ctl, h->T, h->ld, h->m, h->n, e, h->col0,
                                                       h->col1, h->bcol);
  h->next_valid = false;
  h->col_holds = e;
  launch_ratio(h, 1);
  CK(cudaGetLastError());
  int rc = sync_ctl(h);
  if (rc) return rc;
  *l = h->h_ctl->q_leaving;
  return LPS_OK;
}

int lps_pivot(lps_handle h, int e, int l) {
  if (!h) return LPS_ERR_INVALID;
  if (h->sharded) return fail(h, LPS_ERR_STATE, "not available on a row shard (use lps_run)");
  if (!h->loaded) return fail(h, LPS_ERR_STATE, "nothing loaded");
  if (e < 0 || e >= h->n || l < 0 || l >= h->m)
    return fail(h, LPS_ERR_INVALID, "pivot: index out of range");
  CK(cudaSetDevice(h->dev));
  if (h->col_holds != e) {
    k_extract<<<cdiv(h->m + 1, 256), 256, 0, h->stream>>>(h->ctl, h->T, h->ld, h->m, h->n, e,
                                                         h->col0, h->col1, h->bcol);
  }
  k_set_pivot<<<1, 1, 0, h->stream>>>(h->ctl, h->col0, h->col1, e, l, h->n, h->plog, h->log_cap,
                                      h->pos2var);
  launch_scale_update(h, nullptr, nullptr);
  CK(cudaGetLastError());
  int rc = sync_ctl(h);
  if (rc) return rc;
  if (h->h_ctl->status == kZeroPivot) {
    h->next_valid = false;
    h->col_holds = -1;
    return fail(h, LPS_ERR_INVALID, "pivot: pivot element is zero (ArithmeticException in the reference)");
  }
  h->total_pivots = h->h_ctl->npivots;
  h->next_valid = true;  // k_scale_row/k_update staged the next entering column
  h->next_in_nx = false;
  h->col_holds = -2;
  return LPS_OK;
}

int lps_run(lps_handle h, int64_t max_pivots, lps_run_result* res) {
  if (!h) return LPS_ERR_INVALID;
  if (!h->loaded) return fail(h, LPS_ERR_STATE, "nothing loaded");
  CK(cudaSetDevice(h->dev));
  if (use_look(h)) {
    const int rc_look = ensure_look(h);
    if (rc_look == LPS_OK) return run_look(h, max_pivots, res);
    if (h->opt.loop_mode == 7 || rc_look != LPS_ERR_NOMEM) return rc_look;
    // no room for the second tableau buffer: the serial blocked loop runs in place
  }
  if (use_blocked(h)) return run_blocked(h, max_pivots, res);
  const long long start_pivots = h->total_pivots;
  long long launches = 0;
  CK(cudaEventRecord(h->ev_begin, h->stream));
  const bool persistent = use_persistent(h);
  int rc = LPS_OK;
  if (h->next_valid) rc = ensure_next_fmt(h, persistent || h->sharded);
  else rc = (h->sharded || persistent) ? shard_prepare_next(h) : prepare_next(h);
  if (rc) return rc;
  launches += 3;
  if (persistent) {
    ks_begin_run<<<1, 1, 0, h->stream>>>(h->ctls, max_pivots, 0);
    rc = launch_loop(h);
    if (rc) return rc;
    launches += 2;
    rc = sync_ctl(h);
    if (rc) return rc;
    if (h->h_ctl->status == kCommTimeout) return fail(h, LPS_ERR_COMM, "shard: timed out waiting for a peer rank");
    CK(cudaEventRecord(h->ev_end, h->stream));
    CK(cudaEventSynchronize(h->ev_end));
    const long long done = h->h_ctl->npivots - h->total_pivots;
    h->total_pivots = h->h_ctl->npivots;
    h->next_valid = (h->h_ctl->status == kPivotCap);
    h->next_in_nx = true;
    h->col_holds = h->next_valid ? -2 : -1;
    if (res) {
      std::memset(res, 0, sizeof(*res));
      res->verdict = h->h_ctl->status;
      res->last_entering = h->h_ctl->e_cur;
      res->last_leaving = h->h_ctl->l_cur;
      res->npivots = done;
      res->total_pivots = h->total_pivots;
      double corner = 0.0;
      CK(cudaMemcpy(&corner, h->T + (long long)h->m * h->ld + h->n, sizeof(double), cudaMemcpyDeviceToHost));
      res->v = 0.0 - corner;
      cudaEventElapsedTime(&res->device_ms, h->ev_begin, h->ev_end);
      res->update_ms = (float)(h->h_ctls->upd_ns * 1e-6);
      res->update_launches = done;
      res->kernel_launches = launches;
    }
    return LPS_OK;
  }
  if (h->sharded) ks_begin_run<<<1, 1, 0, h->stream>>>(h->ctls, max_pivots, 0);
  else k_begin_run<<<1, 1, 0, h->stream>>>(h->ctl, max_pivots, 0);
  launches++;

  // batch size: about 30 ms of device work per host check, from the tableau's size
  const double bytes = 16.0 * (double)(h->m + 1) * (double)(h->n + 1);
  const double est_us = bytes / 5.0e6 + 12.0;  // ~5 TB/s + fixed launch chain
  long long batch = (long long)(30000.0 / est_us);
  batch = std::max(8ll, std::min(batch, 4096ll));
  const bool timed = h->opt.time_kernels != 0;
  if (timed) {
    while ((long long)h->ev.size() < 2 * batch) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      h->ev.push_back(e);
    }
  }
  double upd_ms = 0.0;
  long long upd_launches = 0;
  long long remaining = (max_pivots < 0) ? -1 : (long long)max_pivots;
  for (;;) {
    // one extra ratio step past the cap is what turns "cap reached" into a verdict
    long long todo = (remaining < 0) ? batch : std::min(batch, remaining + 1);
    for (long long k = 0; k < todo; k++) {
      const bool cap_step = remaining >= 0 && k == remaining;  // the step that only reports PIVOT_CAP
      if (h->sharded) {
        shard_launch_pivot(h, cap_step, timed ? h->ev[2 * k] : nullptr, timed ? h->ev[2 * k + 1] : nullptr);
        launches += cap_step ? 2 : 3;
        if (cap_step) break;
        continue;
      }
      launch_ratio(h, 0);
      launches++;
      if (cap_step) break;
      launch_scale_update(h, timed ? h->ev[2 * k] : nullptr, timed ? h->ev[2 * k + 1] : nullptr);
      launches += 2;
    }
    CK(cudaGetLastError());
    rc = sync_ctl(h);
    if (rc) return rc;
    long long done_now = h->h_ctl->npivots - h->total_pivots;
    if (timed) {
      for (long long k = 0; k < done_now && k < todo; k++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev[2 * k], h->ev[2 * k + 1]) == cudaSuccess) upd_ms += ms;
        upd_launches++;
      }
    }
    h->total_pivots = h->h_ctl->npivots;
    if (remaining >= 0) remaining -= done_now;
    if (h->h_ctl->status != kRunning) break;
  }
  if (h->h_ctl->status == kCommTimeout) return fail(h, LPS_ERR_COMM, "shard: timed out waiting for a peer rank");
  CK(cudaEventRecord(h->ev_end, h->stream));
  CK(cudaEventSynchronize(h->ev_end));
  // after a terminal verdict the staged (e_next, column) pair is still the one the verdict was
  // taken on, so a later call may continue from it (e.g. after PIVOT_CAP)
  h->next_valid = (h->h_ctl->status == kPivotCap);
  h->next_in_nx = h->sharded;
  h->col_holds = h->next_valid ? -2 : -1;
  if (res) {
    std::memset(res, 0, sizeof(*res));
    res->verdict = h->h_ctl->status;
    res->last_entering = h->h_ctl->e_cur;
    res->last_leaving = h->h_ctl->l_cur;
    res->npivots = h->total_pivots - start_pivots;
    res->total_pivots = h->total_pivots;
    double corner = 0.0;
    CK(cudaMemcpy(&corner, h->T + (long long)h->m * h->ld + h->n, sizeof(double), cudaMemcpyDeviceToHost));
    res->v = 0.0 - corner;
    cudaEventElapsedTime(&res->device_ms, h->ev_begin, h->ev_end);
    res->update_ms = (float)upd_ms;
    res->update_launches = upd_launches;
    res->kernel_launches = launches;
  }
  return LPS_OK;
}

int lps_dims(lps_handle h, int* m, int* n) {
  if (!h || !h->loaded) return LPS_ERR_STATE;
  if (m) *m = h->m;
  if (n) *n = h->n;
  return LPS_OK;
}

int lps_read_v(lps_handle h, double* v) {
  if (!h || !v) return LPS_ERR_INVALID;
  if (!h->loaded) return fail(h, LPS_ERR_STATE, "nothing loaded");
  CK(cudaSetDevice(h->dev));
  CK(cudaStreamSynchronize(h->stream));
  double corner = 0.0;
  CK(cudaMemcpy(&corner, h->T + (long long)h->m * h->ld + h->n, sizeof(double), cudaMemcpyDeviceToHost));
  *v = 0.0 - corner;
  return LPS_OK;
}

static int read_column(lps_handle h, int col, double* dst) {
  if (h->m == 0) return LPS_OK;
  k_gather_column<<<cdiv(h->m, 256), 256, 0, h->stream>>>(h->T, h->ld, h->m, col, h->scratch);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(dst, h->scratch, (size_t)h->m * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return LPS_OK;
}

int lps_read_b(lps_handle h, double* b) {
  if (!h || !b) return LPS_ERR_INVALID;
  if (!h->loaded) return fail(h, LPS_ERR_STATE, "nothing loaded");
  CK(cudaSetDevice(h->dev));
  return read_column(h, h->n, b);
}

int lps_read_col(lps_handle h, int j, double* col) {
  if (!h || !col) return LPS_ERR_INVALID;
  if (!h->loaded) return fail(h, LPS_ERR_STATE, "nothing loaded");
  if (j < 0 || j >= h->n) return fail(h, LPS_ERR_INVALID, "read_col: column out of range");
  CK(cudaSetDevice(h->dev));
  return read_column(h, j, col);
}

int lps_read_c(lps_handle h, double* c) {
  if (!h || !c) return LPS_ERR_INVALID;
  if (!h->loaded) return fail(h, LPS_ERR_STATE, "nothing loaded");
  CK(cudaSetDevice(h->dev));
  CK(cudaStreamSynchronize(h->stream));
  if (h->n > 0)
    CK(cudaMemcpy(c, h->T + (long long)h->m * h->ld, (size_t)h->n * sizeof(double), cudaMemcpyDeviceToHost));
  return LPS_OK;
}

int lps_read_row(lps_handle h, int i, double* row) {
  if (!h || !row) return LPS_ERR_INVALID;
  if (!h->loaded) return fail(h, LPS_ERR_STATE, "nothing loaded");
  if (i < 0 || i >= h->m) return fail(h, LPS_ERR_INVALID, "read_row: row out of range");
  CK(cudaSetDevice(h->dev));
  CK(cudaStreamSynchronize(h->stream));
  if (h->n > 0)
    CK(cudaMemcpy(row, h->T + (long long)i * h->ld, (size_t)h->n * sizeof(double), cudaMemcpyDeviceToHost));
  return LPS_OK;
}

int lps_read_A(lps_handle h, double* A, int64_t lda) {
  if (!h || !A || lda < (h ? h->n : 0)) return LPS_ERR_INVALID;
  if (!h->loaded) return fail(h, LPS_ERR_STATE, "nothing loaded");
  CK(cudaSetDevice(h->dev));
  CK(cudaStreamSynchronize(h->stream));
  if (h->m > 0 && h->n > 0)
    CK(cudaMemcpy2D(A, (size_t)lda * sizeof(double), h->T, (size_t)h->ld * sizeof(double),
                    (size_t)h->n * sizeof(double), h->m, cudaMemcpyDeviceToHost));
  return LPS_OK;
}

int lps_read_positions(lps_handle h, int* pos2var) {
  if (!h || !pos2var) return LPS_ERR_INVALID;
  if (!h->loaded) return fail(h, LPS_ERR_STATE, "nothing loaded");
  CK(cudaSetDevice(h->dev));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(pos2var, h->pos2var, (size_t)((h->sharded ? h->m_total : h->m) + h->n) * sizeof(int),
                cudaMemcpyDeviceToHost));
  return LPS_OK;
}

int lps_position_of(lps_handle h, int var, int* pos) {
  if (!h || !pos) return LPS_ERR_INVALID;
  if (!h->loaded) return fail(h, LPS_ERR_STATE, "nothing loaded");
  std::vector<int> p((size_t)(h->sharded ? h->m_total : h->m) + h->n);
  int rc = lps_read_positions(h, p.data());
  if (rc) return rc;
  *pos = -1;
  for (size_t k = 0; k < p.size(); k++)
    if (p[k] == var) { *pos = (int)k; break; }
  return LPS_OK;
}

int lps_read_pivot_log(lps_handle h, int* pairs, int64_t cap_pairs, int64_t* count) {
  if (!h || !count) return LPS_ERR_INVALID;
  if (!h->loaded) return fail(h, LPS_ERR_STATE, "nothing loaded");
  CK(cudaSetDevice(h->dev));
  CK(cudaStreamSynchronize(h->stream));
  long long total = h->total_pivots;
  long long avail = std::min(total, h->log_cap);
  *count = avail;
  if (!pairs || cap_pairs <= 0) return LPS_OK;
  long long take = std::min<long long>(avail, cap_pairs);
  long long first = total - avail;  // oldest retained pivot
  for (long long k = 0; k < take;) {
    long long slot = (first + k) % h->log_cap;
    long long run = std::min(take - k, h->log_cap - slot);
    CK(cudaMemcpy(pairs + 2 * k, h->plog + slot, (size_t)run * sizeof(int2), cudaMemcpyDeviceToHost));
    k += run;
  }
  return LPS_OK;
}

int lps_read_primal(lps_handle h, int nvars, double* x) {
  if (!h || !x || nvars < 0) return LPS_ERR_INVALID;
  if (!h->loaded) return fail(h, LPS_ERR_STATE, "nothing loaded");
  if (h->sharded) return fail(h, LPS_ERR_STATE, "read_primal: gather b from the shards first");
  std::vector<int> p((size_t)h->m + h->n);
  std::vector<double> b((size_t)h->m);
  int rc = lps_read_positions(h, p.data());
  if (rc) return rc;
  rc = lps_read_b(h, b.data());
  if (rc) return rc;
  for (int k = 0; k < nvars; k++) x[k] = 0.0;
  for (int pos = h->n; pos < h->n + h->m; pos++) {
    int var = p[pos];
    if (var >= 0 && var < nvars) x[var] = b[pos - h->n];
  }
  return LPS_OK;
}

int lps_first_nonzero_in_row(lps_handle h, int row, int* j) {
  if (!h || !j) return LPS_ERR_INVALID;
  if (h->sharded) return fail(h, LPS_ERR_STATE, "not available on a row shard (use lps_run)");
  if (!h->loaded) return fail(h, LPS_ERR_STATE, "nothing loaded");
  if (row < 0 || row >= h->m) return fail(h, LPS_ERR_INVALID, "first_nonzero_in_row: row out of range");
  CK(cudaSetDevice(h->dev));
  const int none = kNone;
  CK(cudaMemcpyAsync(&h->ctl->q_index, &none, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  k_first_nonzero<<<cdiv(h->n, 256), 256, 0, h->stream>>>(h->ctl, h->T + (long long)row * h->ld, h->n,
                                                         h->opt.epsilon);
  CK(cudaGetLastError());
  int rc = sync_ctl(h);
  if (rc) return rc;
  *j = (h->h_ctl->q_index == kNone) ? -1 : h->h_ctl->q_index;
  return LPS_OK;
}

int lps_drop_column(lps_handle h, int j) {
  if (!h) return LPS_ERR_INVALID;
  if (h->sharded) return fail(h, LPS_ERR_STATE, "not available on a row shard (use lps_run)");
  if (!h->loaded) return fail(h, LPS_ERR_STATE, "nothing loaded");
  if (j < 0 || j >= h->n) return fail(h, LPS_ERR_INVALID, "drop_column: column out of range");
  CK(cudaSetDevice(h->dev));
  const int threads = 256;
  int grid = std::min(h->m + 1, 8 * h->sm_count);
  k_drop_column<<<grid, threads, 0, h->stream>>>(h->T, h->ld, h->m, h->n, j);
  CK(cudaGetLastError());
  // positions j+1.. shift down by one (LPSolver.java:239-244)
  std::vector<int> p((size_t)h->m + h->n);
  CK(cudaMemcpyAsync(p.data(), h->pos2var, p.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  p.erase(p.begin() + j);
  // on the handle's own (non-blocking) stream and waited for: later kernels on it swap entries of the map
  CK(cudaMemcpyAsync(h->pos2var, p.data(), p.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->n -= 1;
  h->next_valid = false;
  h->col_holds = -1;
  return LPS_OK;
}

int lps_rebuild_objective(lps_handle h, const lps_objective_op* ops, int nops) {
  if (!h || nops < 0 || (nops > 0 && !ops)) return LPS_ERR_INVALID;
  if (!h->loaded) return fail(h, LPS_ERR_STATE, "nothing loaded");
  for (int k = 0; k < nops; k++) {
    if (ops[k].kind == 0 ? (ops[k].index < 0 || ops[k].index >= h->m)
                         : (ops[k].kind != 1 || ops[k].index < 0 || ops[k].index >= h->n))
      return fail(h, LPS_ERR_INVALID, "rebuild_objective: op index out of range");
  }
  CK(cudaSetDevice(h->dev));
  if ((size_t)nops > h->ops_cap) {
    if (h->d_ops) cudaFree(h->d_ops);
    h->ops_cap = (size_t)nops + 64;
    CK(cudaMalloc(&h->d_ops, h->ops_cap * sizeof(ObjOp)));
  }
  static_assert(sizeof(ObjOp) == sizeof(lps_objective_op), "op layout");
  if (nops > 0)
    CK(cudaMemcpyAsync(h->d_ops, ops, (size_t)nops * sizeof(ObjOp), cudaMemcpyHostToDevice, h->stream));
  k_rebuild_objective<<<cdiv(h->n + 1, 128), 128, 0, h->stream>>>(h->T, h->ld, h->m, h->n, h->d_ops, nops);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  h->next_valid = false;
  h->col_holds = -1;
  return LPS_OK;
}

int lps_device_info(lps_handle h, int* sm_count, int64_t* hbm_bytes, int* cc_major, int* cc_minor) {
  if (!h) return LPS_ERR_INVALID;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, h->dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (hbm_bytes) *hbm_bytes = (int64_t)prop.totalGlobalMem;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return LPS_OK;
}

int lps_tableau_bytes(lps_handle h, int64_t* bytes) {
  if (!h || !bytes || !h->loaded) return LPS_ERR_STATE;
  *bytes = (int64_t)sizeof(double) * (h->m + 1) * h->ld;
  return LPS_OK;
}

int lps_algorithmic_bytes_per_pivot(lps_handle h, int64_t* bytes) {
  if (!h || !bytes || !h->loaded) return LPS_ERR_STATE;
  *bytes = 16ll * (h->m + 1) * (h->n + 1);
  return LPS_OK;
}

int lps_measure_fp64_issue_rate(lps_handle h, double ms, double* inst_per_s) {
  if (!h || !inst_per_s) return LPS_ERR_INVALID;
  CK(cudaSetDevice(h->dev));
  constexpr int kChains = 16;
  const int grid = 4 * h->sm_count, threads = 256, iters = 2048;
  double* out = reinterpret_cast<double*>(h->partials);
  k_fp64_probe<kChains><<<grid, threads, 0, h->stream>>>(out, 0.5, 16);     // warm-up
  const double per_launch = (double)grid * threads * iters * kChains * 2.0;
  // about 0.15 ms per launch at 18 T inst/s: enough launches to fill `ms`
  const int reps = std::max(1, std::min(4096, (int)(ms / 0.15)));
  CK(cudaEventRecord(h->ev_begin, h->stream));
  for (int k = 0; k < reps; k++) k_fp64_probe<kChains><<<grid, threads, 0, h->stream>>>(out, 0.5, iters);
  CK(cudaEventRecord(h->ev_end, h->stream));
  CK(cudaEventSynchronize(h->ev_end));
  CK(cudaGetLastError());
  float el = 0.f;
  CK(cudaEventElapsedTime(&el, h->ev_begin, h->ev_end));
  *inst_per_s = el > 0.f ? per_launch * reps / (el * 1e-3) : 0.0;
  return LPS_OK;
}

int lps_plan_split_model(int grid, int block_pivots, int world, int64_t rows_local, int64_t pitch) {
  if (grid < 2 || block_pivots < 1 || world < 1 || rows_local < 0 || pitch < 1) return LPS_ERR_INVALID;
  return split_model(grid, block_pivots, world, rows_local, pitch);
}
int lps_plan_split_tuned(int grid, int block_pivots, int world, int current, double panel_us_per_pivot,
                         double pass_us_per_block) {
  if (grid < 2 || block_pivots < 1 || world < 1) return LPS_ERR_INVALID;
  return split_tuned(grid, block_pivots, world, current, panel_us_per_pivot, pass_us_per_block);
}
int lps_loop_description(lps_handle h, char* buf, int cap) {
  if (!h || !buf || cap <= 0) return LPS_ERR_INVALID;
  if (!h->loaded) return fail(h, LPS_ERR_STATE, "nothing loaded");
  char tmp[512];
  if (use_look(h)) {
    const int P = look_panel_ctas(h);
    if (step_is_ws(h))
      std::snprintf(tmp, sizeof tmp, "look-ahead blocked loop, %d pivots per pass: lps::kb_step_ws (12 pass warps + 4 panel warps in each of %d CTAs)",
                    h->block, h->sm_count);
    else if (step_uses_flush(h))
      std::snprintf(tmp, sizeof tmp, "look-ahead blocked loop, %d pivots per pass: lps::kb_step_flush (panel role on %d CTAs, cp.async pass role on %d)",
                    h->block, P, h->sm_count - P);
    else
      std::snprintf(tmp, sizeof tmp, "look-ahead blocked loop, %d pivots per pass: lps::kb_step (panel role on %d CTAs, TMA + mbarrier pass role on %d, %d-row stages)",
                    h->block, P, h->sm_count - P, pass_stage_rows(h));
  } else if (use_blocked(h)) {
    std::snprintf(tmp, sizeof tmp, "serial blocked loop, %d pivots per pass: lps::%s then lps::%s", h->block,
                  h->opt.loop_mode == 5 ? "kb_col + kb_row per pivot" : "kb_panel",
                  (h->opt.update_variant >= 10 && sweep_available(h)) ? "kb_sweep (TMA)" : "kb_flush (cp.async)");
  } else if (use_persistent(h)) {
    std::snprintf(tmp, sizeof tmp, "one pass per pivot, persistent cooperative loop: lps::k_loop");
  } else {
    std::snprintf(tmp, sizeof tmp, "one pass per pivot: lps::%s", h->sharded ? "ks_ratio, ks_scale_row, ks_update" : "k_ratio, k_scale_row, k_update");
  }
  std::snprintf(buf, (size_t)cap, "%s", tmp);
  return LPS_OK;
}

// ---- row-sharded API ------------------------------------------------------------------------
int lps_shard_generate_lp(lps_handle h, int kind, int m_total, int n, int rank, int world, uint64_t seed,
                          int param) {
  if (!h || m_total <= 0 || n <= 0) return fail(h, LPS_ERR_INVALID, "lps_shard_generate_lp: bad dimensions");
  if (kind < 0 || kind > 2 || (kind == LPS_GEN_UNBOUNDED && (param < 0 || param >= n)) ||
      (kind == LPS_GEN_ASSIGNMENT && m_total < 2))
    return fail(h, LPS_ERR_INVALID, "lps_shard_generate_lp: bad kind / parameter");
  CK(cudaSetDevice(h->dev));
  int rc = shard_setup(h, m_total, n, rank, world);
  if (rc) return rc;
  dim3 grid(std::min(cdiv(h->ld, 256), 64), std::min(h->m + 1, 65535));
  ks_generate_lp<<<grid, 256, 0, h->stream>>>(h->T, h->ld, m_total, n, h->row0, h->row1, seed, kind, param);
  CK(cudaGetLastError());
  rc = shard_reset_state(h);
  if (rc) return rc;
  CK(cudaStreamSynchronize(h->stream));
  return LPS_OK;
}

int lps_shard_generate_dense(lps_handle h, int m_total, int n, int rank, int world, uint64_t seed,
                             int pos_permille) {
  return lps_shard_generate_lp(h, LPS_GEN_DENSE, m_total, n, rank, world, seed, pos_permille);
}

int lps_shard_load(lps_handle h, int m_total, int n, int rank, int world, const double* A_local,
                   int64_t lda, const double* b_local, const double* c, double v) {
  if (!h) return LPS_ERR_INVALID;
  if (lda < n || (n > 0 && !c)) return fail(h, LPS_ERR_INVALID, "lps_shard_load: bad arguments");
  CK(cudaSetDevice(h->dev));
  int rc = shard_setup(h, m_total, n, rank, world);
  if (rc) return rc;
  const int m = h->m;
  const long long ld = h->ld;
  if (m > 0 && (!A_local || !b_local)) return fail(h, LPS_ERR_INVALID, "lps_shard_load: null buffer");
  CK(cudaMemsetAsync(h->T, 0, (size_t)(m + 1) * ld * sizeof(double), h->stream));
  if (m > 0 && n > 0)
    CK(cudaMemcpy2DAsync(h->T, ld * sizeof(double), A_local, (size_t)lda * sizeof(double),
                         (size_t)n * sizeof(double), m, cudaMemcpyHostToDevice, h->stream));
  if (m > 0) {
    CK(cudaMemcpyAsync(h->scratch, b_local, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    k_set_column<<<cdiv(m, 256), 256, 0, h->stream>>>(h->T, ld, m, n, h->scratch, 0.0, 0);
  }
  if (n > 0)
    CK(cudaMemcpyAsync(h->T + (long long)m * ld, c, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  const double negv = -v;
  CK(cudaMemcpyAsync(h->T + (long long)m * ld + n, &negv, sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaGetLastError());
  rc = shard_reset_state(h);
  if (rc) return rc;
  CK(cudaStreamSynchronize(h->stream));
  return LPS_OK;
}

int lps_shard_info(lps_handle h, int* rank, int* world, int* m_total, int* row0, int* row1) {
  if (!h || !h->sharded) return LPS_ERR_STATE;
  if (rank) *rank = h->rank;
  if (world) *world = h->world;
  if (m_total) *m_total = h->m_total;
  if (row0) *row0 = h->row0;
  if (row1) *row1 = h->row1;
  return LPS_OK;
}

int lps_shard_export(lps_handle h, void* handle64) {
  if (!h || !handle64) return LPS_ERR_INVALID;
  if (!h->sharded || !h->comm) return fail(h, LPS_ERR_STATE, "shard: nothing loaded");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  CK(cudaSetDevice(h->dev));
  cudaIpcMemHandle_t ih;
  CK(cudaIpcGetMemHandle(&ih, h->comm));
  std::memcpy(handle64, &ih, 64);
  return LPS_OK;
}

int lps_shard_comm_ptr(lps_handle h, void** p) {
  if (!h || !p) return LPS_ERR_INVALID;
  if (!h->sharded || !h->comm) return fail(h, LPS_ERR_STATE, "shard: nothing loaded");
  *p = h->comm;
  return LPS_OK;
}

static int attach_common(lps_handle h, void* const* blocks) {
  for (int k = 0; k < h->world; k++) {
    CommBlock* b = (k == h->rank) ? h->comm : reinterpret_cast<CommBlock*>(blocks[k]);
    if (!b) return fail(h, LPS_ERR_INVALID, "shard attach: null peer block");
    h->peers.blk[k] = b;
    h->peers.rowbuf[k] = reinterpret_cast<double*>(reinterpret_cast<char*>(b) + sizeof(CommBlock));
  }
  h->attached = true;
  return LPS_OK;
}

int lps_shard_attach_ptrs(lps_handle h, void* const* comm_ptrs) {
  if (!h || !comm_ptrs) return LPS_ERR_INVALID;
  if (!h->sharded || !h->comm) return fail(h, LPS_ERR_STATE, "shard: nothing loaded");
  CK(cudaSetDevice(h->dev));
  // same-process peers: make their memory addressable from this device
  for (int k = 0; k < h->world; k++) {
    if (k == h->rank || !comm_ptrs[k]) continue;
    cudaPointerAttributes at;
    CK(cudaPointerGetAttributes(&at, comm_ptrs[k]));
    if (at.device != h->dev) {
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, h->dev, at.device));
      if (!can) return fail(h, LPS_ERR_COMM, "shard attach: no peer access between the two GPUs");
      cudaError_t ce = cudaDeviceEnablePeerAccess(at.device, 0);
      if (ce != cudaSuccess && ce != cudaErrorPeerAccessAlreadyEnabled) return fail(h, LPS_ERR_COMM, "cudaDeviceEnablePeerAccess", ce);
      cudaGetLastError();
    }
  }
  return attach_common(h, comm_ptrs);
}

int lps_shard_attach_ipc(lps_handle h, const void* handles /* world x 64 bytes, rank order */) {
  if (!h || !handles) return LPS_ERR_INVALID;
  if (!h->sharded || !h->comm) return fail(h, LPS_ERR_STATE, "shard: nothing loaded");
  CK(cudaSetDevice(h->dev));
  void* blocks[kMaxRanks] = {nullptr};
  for (int k = 0; k < h->world; k++) {
    if (k == h->rank) continue;
    cudaIpcMemHandle_t ih;
    std::memcpy(&ih, static_cast<const char*>(handles) + 64 * k, 64);
    void* p = nullptr;
    cudaError_t ce = cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess);
    if (ce != cudaSuccess) return fail(h, LPS_ERR_COMM, "cudaIpcOpenMemHandle", ce);
    h->ipc_opened.push_back(p);
    blocks[k] = p;
  }
  return attach_common(h, blocks);
}

}  // extern "C"
