/*
 * lps_b200.h — C ABI of the B200-native dense-tableau simplex pivot loop.
 *
 * Drop-in boundary: the reference (Toptachamann/Linear_Programming_Solver) has no FFI seam;
 * its hot path is the Java class `lpsolver.LPState` and the two loops in `lpsolver.LPSolver`
 * that drive it.  Each entry point below replaces one of those members — the reference
 * location is cited as file:line under src/main/java/lpsolver/.  A JNI / Panama binding calls
 * exactly these symbols.  Plain C types only; no exception crosses the
 * boundary; every function returns an `lps_status` (0 = ok, < 0 = error) and verdicts are data.
 *
 * The tableau lives in HBM for the life of the handle as the augmented matrix
 *     T = [ A | b ]   (m rows)
 *         [ c | -v ]  (1 row)
 * in binary64, row-major with a padded pitch.  Arithmetic is IEEE-754 round-to-nearest with a
 * separately rounded multiply and subtract per cell (no FMA) and true division, mirroring the
 * reference's one-rounding-per-BigDecimal-operation structure (LPState.java:139-177).
 *
 * Thread safety: a handle is single-caller (like LPState); distinct handles are independent.
 */
#ifndef LPS_B200_H
#define LPS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LPS_ABI_VERSION 1

typedef struct lps_handle_s *lps_handle;

/* return codes */
typedef enum {
  LPS_OK = 0,
  LPS_ERR_INVALID = -1,   /* bad argument (range errors on e / l: Validate.isTrue, LPState.java:288) */
  LPS_ERR_CUDA = -2,      /* CUDA runtime failure; text in lps_last_error */
  LPS_ERR_STATE = -3,     /* call made in the wrong state (e.g. nothing loaded) */
  LPS_ERR_NOMEM = -4,
  LPS_ERR_NODEVICE = -5,  /* no usable CUDA device: there is no CPU fallback */
  LPS_ERR_COMM = -6       /* multi-GPU exchange failure */
} lps_status;

/* verdict of lps_run (data, not errors) */
typedef enum {
  LPS_RUNNING = 0,
  LPS_OPTIMAL = 1,    /* getEntering() == -1                      LPSolver.java:101 */
  LPS_UNBOUNDED = 2,  /* getLeaving(e) == -1                      LPSolver.java:103-106, :147-150 */
  LPS_PIVOT_CAP = 3   /* max_pivots reached with a pivot pending  (no reference analogue) */
} lps_verdict;

typedef struct {
  double epsilon;       /* LPState.DEF_EPSILON = 1e-9   LPState.java:20 */
  double inf;           /* LPState.DEF_INF     = 1e50   LPState.java:21 */
  int device;           /* CUDA device ordinal, -1 = current device */
  int time_kernels;     /* != 0: bracket every tableau-update launch with CUDA events */
  void *stream;         /* cudaStream_t to run on, NULL = the handle creates its own */
  int update_variant;   /* tableau-update kernel: -1 = default, >= 0 selects an alternative (tuning) */
  int loop_mode;        /* lps_run: 0 = auto, 1 = three kernels per pivot, 2 = one persistent
                           cooperative kernel for the whole loop (grid barriers between phases),
                           5 = blocked loop (block_pivots pivots per tableau pass) with two launches per
                           pivot, 6 = blocked loop with one cooperative panel launch per block followed by
                           the pass, 7 = look-ahead blocked loop: ONE cooperative launch per block runs the
                           pass of block k (out of place; TMA pipeline on shards >= 2.5 GB, cp.async pass
                           below, update_variant >= 10 / 0..9 force either) and the panel of block k+1 side by
                           side on disjoint SMs (needs a second tableau buffer; the default above 64 MB), 8 = the same
                           with the pass warps and the panel warps inside every CTA (kb_step_ws) */
  int block_pivots;     /* lps_run: pivots deferred between two passes over the tableau (blocked loop):
                           0 = default (16), 1 = off (every pivot is its own pass), at most 20 (16 for
                           loop_mode 7).  Values are bit-identical for every setting. */
  int panel_ctas;       /* loop_mode 7: CTAs (= SMs) given to the panel role, the pass gets the rest; 0 = auto:
                           lps_plan_split_model for a handle's first run, then re-fitted after every run from
                           the two roles' measured durations (lps_plan_split_tuned) */
  int pass_chunk_rows;  /* TMA pass: rows per work chunk (rounded to a multiple of 12), 0 = auto */
  int reserved[3];
} lps_options;

typedef struct {
  int verdict;          /* lps_verdict */
  int last_entering;    /* entering index of the last pivot considered (-1 if none) */
  int last_leaving;
  int pad_;
  int64_t npivots;      /* pivots executed by this call */
  int64_t total_pivots; /* pivots executed on this handle since the last load */
  double v;             /* objective value after the call (LPState.v) */
  float device_ms;      /* CUDA-event time of the whole call on the handle's stream */
  float update_ms;      /* sum of tableau-update kernel times (time_kernels only) */
  int64_t update_launches;
  int64_t kernel_launches; /* every kernel this call launched */
} lps_run_result;

/* one step of LPSolver.restoreInitialLP's objective rebuild (LPSolver.java:217-233) */
typedef struct {
  int kind;     /* 0: basic variable  -> v += b[index]*coef ; c[j] += (-A[index][j])*coef  (:223-227)
                   1: non-basic       -> c[index] += coef                                   (:231) */
  int index;    /* row (kind 0) or column (kind 1) in the tableau AFTER the column drop */
  double coef;  /* the variable's coefficient in the initial objective */
} lps_objective_op;

int lps_abi_version(void);
void lps_default_options(lps_options *opts);
const char *lps_status_string(int status);

/* life cycle ------------------------------------------------------------------------------ */
int lps_create(lps_handle *out, const lps_options *opts);
int lps_destroy(lps_handle h);
const char *lps_last_error(lps_handle h); /* valid until the next call on h */

/* LPState(A, b, c, v, …, m, n)  — LPState.java:37-112.  Host buffers are COPIED to HBM (the
 * reference aliases them, LPSolver.java:267,270; the shim writes results back if it wants the
 * aliasing).  A is row-major with leading dimension lda >= n. */
int lps_load(lps_handle h, int m, int n, const double *A, int64_t lda, const double *b,
             const double *c, double v);

/* LPSolver.convertIntoAuxLP — LPSolver.java:283-321: loads A into an m x (n+1) tableau whose
 * last column is -1, objective (0,…,0,-1), v = 0.  x0 is variable id n at position n. */
int lps_load_aux(lps_handle h, int m, int n, const double *A, int64_t lda, const double *b);

/* synthetic dense LP generated directly in HBM (bench / tests):
 * A_ij = u(i*n+j), c_j = ±u(mn+j), b_i = (n/4)(1+u(mn+n+i)).  rows [row0,row1) only are
 * materialised when the handle is a row shard (row0=0,row1=m for a single GPU). */
int lps_generate_dense(lps_handle h, int m, int n, uint64_t seed, int pos_permille);
/* the other synthetic families of SURVEY.md §8d, same generator:
 *   LPS_GEN_DENSE       as lps_generate_dense (param = pos_permille)
 *   LPS_GEN_UNBOUNDED   the dense LP with every c_j > 0 and column `param` of A negated (no positive
 *                       entry): unbounded, reported when that column enters (immediately for param = 0)
 *   LPS_GEN_ASSIGNMENT  degenerate: 0/1 incidence matrix of a bipartite graph, b = 1, c = 1; totally
 *                       unimodular, so every entry stays in {-1,0,1} (exact in binary64 and in the
 *                       reference's decimal arithmetic) and zero-ratio ties abound */
typedef enum { LPS_GEN_DENSE = 0, LPS_GEN_UNBOUNDED = 1, LPS_GEN_ASSIGNMENT = 2 } lps_gen_kind;
int lps_generate_lp(lps_handle h, int kind, int m, int n, uint64_t seed, int param);

/* LPState.getEntering() — LPState.java:274-285.  *e = -1 when no c[i] > epsilon. */
int lps_get_entering(lps_handle h, int *e);
/* LPState.getLeaving(int) — LPState.java:287-305.  *l = -1 when no ratio < INF. */
int lps_get_leaving(lps_handle h, int e, int *l);
/* LPState.pivot(int,int) — LPState.java:114-181 (and exchangeIndexes :311-320). */
int lps_pivot(lps_handle h, int e, int l);

/* The loops of LPSolver.simplex (LPSolver.java:101-112) and LPSolver.solveAuxLP (:141-161)
 * run on the device: getEntering / getLeaving / pivot until optimal, unbounded or max_pivots
 * (< 0 = unlimited) further pivots.  The host is not in the per-pivot path. */
int lps_run(lps_handle h, int64_t max_pivots, lps_run_result *res);

/* field reads (LPSolver reads LPState.A/b/c/v directly: LPSolver.java:113,170,185,203-245) */
int lps_dims(lps_handle h, int *m, int *n);
int lps_read_v(lps_handle h, double *v);
int lps_read_b(lps_handle h, double *b);            /* m values */
int lps_read_c(lps_handle h, double *c);            /* n values */
int lps_read_row(lps_handle h, int i, double *row); /* n values of A[i] */
int lps_read_col(lps_handle h, int j, double *col); /* m values of A[.][j] */
int lps_read_A(lps_handle h, double *A, int64_t lda);
/* variables / coefficients maps (LPState.java:27-28) as a permutation: pos2var[pos] = id of
 * the variable at position pos (0..n-1 non-basic columns, n..n+m-1 basic rows). */
int lps_read_positions(lps_handle h, int *pos2var);
int lps_position_of(lps_handle h, int var, int *pos);
/* (entering, leaving) pairs since the last load, oldest first; *count = pairs available */
int lps_read_pivot_log(lps_handle h, int *pairs, int64_t cap_pairs, int64_t *count);
/* primal values of variables 0..nvars-1: b[pos-n] if basic else 0 (io_files/output.txt:214-233) */
int lps_read_primal(lps_handle h, int nvars, double *x);

/* phase-1 support (LPSolver.java:166-246) */
/* performDegeneratePivot's scan — LPSolver.java:185-191: first j with |A[row][j]| > epsilon */
int lps_first_nonzero_in_row(lps_handle h, int row, int *j);
/* restoreInitialLP — LPSolver.java:205-211,235-244: remove column j (positions shift down) */
int lps_drop_column(lps_handle h, int j);
/* restoreInitialLP — LPSolver.java:213-233: c <- 0, v <- 0, then apply ops in order */
int lps_rebuild_objective(lps_handle h, const lps_objective_op *ops, int nops);

/* ---- row sharding across the GPUs of one NVSwitch box -------------------------
 * Rank k of `world` (<= 8) owns rows [k*m/world, (k+1)*m/world) of (A | b) — the block split of
 * LPState.pivotConcurrently (LPState.java:222-223) — plus a replica of the objective row.  One
 * handle per GPU; one host process per GPU (multi-process, CUDA IPC) or one process driving
 * several handles (peer pointers).  Per pivot the ranks exchange their ratio-test candidates
 * and the owner of the leaving row broadcasts the scaled pivot row by direct stores into peer
 * memory over NVLink, inside the kernels; lps_run must then be called on every rank with the
 * same max_pivots.  Field reads return the LOCAL rows; positions and the pivot log are global. */
int lps_shard_generate_dense(lps_handle h, int m_total, int n, int rank, int world, uint64_t seed,
                             int pos_permille);
int lps_shard_generate_lp(lps_handle h, int kind, int m_total, int n, int rank, int world, uint64_t seed,
                          int param);
int lps_shard_load(lps_handle h, int m_total, int n, int rank, int world, const double *A_local,
                   int64_t lda, const double *b_local, const double *c, double v);
int lps_shard_info(lps_handle h, int *rank, int *world, int *m_total, int *row0, int *row1);
/* exchange block: export as a 64-byte CUDA IPC handle (multi-process) or a raw device pointer
 * (same process, peer access enabled by the caller), then attach all ranks' blocks in rank order */
int lps_shard_export(lps_handle h, void *handle64);
int lps_shard_comm_ptr(lps_handle h, void **ptr);
int lps_shard_attach_ipc(lps_handle h, const void *handles);
int lps_shard_attach_ptrs(lps_handle h, void *const *comm_ptrs);

/* introspection */
int lps_device_info(lps_handle h, int *sm_count, int64_t *hbm_bytes, int *cc_major, int *cc_minor);
int lps_tableau_bytes(lps_handle h, int64_t *bytes);     /* 8*(m+1)*pitch */
int lps_algorithmic_bytes_per_pivot(lps_handle h, int64_t *bytes); /* 16*(m+1)*(n+1) */
/* Measurement aid for the roofline of the blocked pass (no reference analogue): issue rate of the pass's
   own instruction mix -- separately rounded DMUL + DADD on independent chains, 256 threads x 4 CTAs per SM --
   in FP64 thread-instructions per second, timed with CUDA events on the handle's stream for about `ms`
   milliseconds.  This is the FP64 roof the pass is held against beside the HBM one. */
int lps_measure_fp64_issue_rate(lps_handle h, double ms, double *inst_per_s);
/* Which loop shape lps_run uses for the loaded LP (bench / log records): kernel names, SM split, pivots per pass. */
int lps_loop_description(lps_handle h, char *buf, int cap);
/* The SM split of the look-ahead step (loop_mode 7), as pure host arithmetic (no device, no handle; no reference
   analogue): how many of `grid` CTAs decide the next block's pivots while the others apply the current block.
   lps_plan_split_model: the a-priori choice for a shard of rows_local x pitch doubles on `world` ranks -- what the
   first run of a handle uses.  lps_plan_split_tuned: the choice after a run that measured the panel role at
   panel_us_per_pivot on `current` CTAs and the pass role at pass_us_per_block on the others -- what every later
   run uses (options.panel_ctas > 0 overrides both).  Both return the CTA count (>= 1), or LPS_ERR_INVALID.
   The split changes timing only: pivots and results are identical for every value. */
int lps_plan_split_model(int grid, int block_pivots, int world, int64_t rows_local, int64_t pitch);
int lps_plan_split_tuned(int grid, int block_pivots, int world, int current, double panel_us_per_pivot,
                         double pass_us_per_block);

#ifdef __cplusplus
}
#endif
#endif /* LPS_B200_H */
