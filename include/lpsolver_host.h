/*
 * lpsolver_host.h — host side of the drop-in: a C++ mirror of the reference's `lpsolver`
 * package surface for the simplex path, written over the C ABI of lps_b200.h.
 *
 * The reference's host language is Java and no JDK exists in the build image, so the host
 * layer above the C ABI is C++ (the reference is compiled code).  Class and method names,
 * argument meaning and error behaviour follow the Java classes so that tests read like the
 * reference's own Spock specs:
 *
 *   lpsolver::LPStandardForm     LPStandardForm.java:10-65
 *   lpsolver::LPState            LPState.java:17-320     (tableau resident in HBM)
 *   lpsolver::LPSolver           LPSolver.java:15-401    (phase-1/phase-2 orchestration)
 *   lpsolver::LPException        LPException.java
 *   lpsolver::SolutionException  SolutionException.java  (extends LPException)
 *
 * A flat C entry point, lpsolver_solve(), wraps LPSolver::solve for FFI callers (ctypes in the
 * tests; a JNI / Panama caller would normally bind lps_b200.h directly).
 */
#ifndef LPSOLVER_HOST_H
#define LPSOLVER_HOST_H

#include "lps_b200.h"

#ifdef __cplusplus
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace lpsolver {

class LPException : public std::runtime_error {
 public:
  explicit LPException(const std::string& m) : std::runtime_error(m) {}
};
class SolutionException : public LPException {
 public:
  explicit SolutionException(const std::string& m) : LPException(m) {}
};
/* java.lang.ArrayIndexOutOfBoundsException escaping restoreInitialLP (LPSolver.java:231) */
class ArrayIndexOutOfBounds : public std::out_of_range {
 public:
  explicit ArrayIndexOutOfBounds(const std::string& m) : std::out_of_range(m) {}
};

struct LPStandardForm {
  int m = 0, n = 0;
  bool maximize = true;
  const double* A = nullptr; /* row-major m x n view with leading dimension lda (not owned) */
  int64_t lda = 0;
  std::vector<double> b, c;
  std::vector<double> A_storage; /* optional owner of A */
  void setA(std::vector<double> a) { A_storage = std::move(a); A = A_storage.data(); lda = n; }
};

/* LPState: the tableau lives on the GPU; fields are read on demand. */
class LPState {
 public:
  LPState(const double* A, int64_t lda, const double* b, const double* c, double v, int m, int n,
          const lps_options& opts);
  /* auxiliary-LP constructor (LPSolver.convertIntoAuxLP) */
  static LPState aux(const double* A, int64_t lda, const double* b, int m, int n,
                     const lps_options& opts);
  ~LPState();
  LPState(LPState&& o) noexcept;
  LPState(const LPState&) = delete;
  LPState& operator=(const LPState&) = delete;

  int getEntering();
  int getLeaving(int entering);
  void pivot(int entering, int leaving);
  lps_run_result run(int64_t max_pivots);

  int m() const;
  int n() const;
  double v();
  std::vector<double> b();
  std::vector<double> c();
  std::vector<double> row(int i);
  std::vector<int> positions();
  int positionOf(int var);
  std::vector<std::pair<int, int>> pivotLog();
  lps_handle handle() { return h_; }
  lps_handle release() { lps_handle t = h_; h_ = nullptr; return t; }

 private:
  explicit LPState(lps_handle h) : h_(h) {}
  void check(int rc, const char* what);
  lps_handle h_ = nullptr;
};

struct SolveTrace {
  bool used_phase1 = false;
  int x0_index = -1;
  int64_t phase1_pivots = 0, phase2_pivots = 0;
  std::vector<std::pair<int, int>> phase1_log, phase2_log;
  double raw_v = 0.0;
  float device_ms = 0.f;
  int verdict = 0;
};

class LPSolver {
 public:
  LPSolver();
  /* epsilon / inf: LPSolver(printRounder, rounder, epsilon, inf), LPSolver.java:45-58.  The
   * MathContext arguments have no binary64 analogue (IEEE round-to-nearest is used). */
  LPSolver(double epsilon, double inf);
  /* returns LPState.v (negated for min); `solve6` applies setScale(6, HALF_UP) like :113 */
  double solve(LPStandardForm& form);
  static std::string setScale6(double v);
  static int minInB(const std::vector<double>& b);

  lps_options options;
  bool fix_restore_index = false; /* false = as the reference is written (LPSolver.java:220,231) */
  int64_t max_pivots = -1;        /* cap on phase-2 pivots (-1 = none; no reference analogue) */
  SolveTrace trace;
  std::vector<double> primal;     /* filled after an optimal solve */
  lps_handle keep_state = nullptr; /* final LPState's handle when keep is set; caller destroys */
  bool keep = false;

 private:
  LPState initializeSimplex(LPStandardForm& form);
};

}  // namespace lpsolver
extern "C" {
#endif /* __cplusplus */

enum {
  LPSOLVER_OPTIMAL = 0,
  LPSOLVER_UNBOUNDED = 1,        /* SolutionException("This linear program is unbounded") */
  LPSOLVER_INFEASIBLE = 2,       /* LPException("This linear program is infeasible") */
  LPSOLVER_AUX_UNBOUNDED = 3,    /* SolutionException("Auxiliary lp is unbounded") */
  LPSOLVER_DEGENERATE_FAIL = 4,  /* SolutionException("Can't perform degenerate pivot") */
  LPSOLVER_INDEX_ERROR = 5,      /* ArrayIndexOutOfBoundsException out of restoreInitialLP */
  LPSOLVER_PIVOT_CAP = 6,
  LPSOLVER_ERROR = 7             /* CUDA / argument failure; message holds lps_last_error */
};

typedef struct {
  int verdict;
  int used_phase1;
  int x0_index;
  int pad_;
  int64_t phase1_pivots;
  int64_t phase2_pivots;
  double value;        /* raw objective (sign restored for min), before 6-decimal rounding */
  float device_ms;
  float pad2_;
  char value6[48];     /* BigDecimal.setScale(6, HALF_UP).toString(), LPSolver.java:113 */
  char message[160];   /* exception text, verbatim where the reference has one */
} lpsolver_result;

/* LPSolver(…).solve(LPStandardForm) — LPSolver.java:78-94.  `c` is negated IN PLACE for a
 * minimisation, as the reference does (:86-89).  primal (n values) and the two pivot logs
 * ((entering, leaving) int pairs) are optional outputs.  If keep_state is non-NULL it receives
 * the handle of the final LPState (caller must lps_destroy it). */
int lpsolver_solve(const lps_options* opts, int m, int n, const double* A, int64_t lda,
                   const double* b, double* c, int maximize, int fix_restore_index,
                   int64_t max_pivots, lpsolver_result* res, double* primal, int* phase1_log,
                   int64_t phase1_cap, int* phase2_log, int64_t phase2_cap, lps_handle* keep_state);

/* LPInputReader.readLP — LPInputReader.java:52-114, the reference's text grammar (:25-31) with the
 * `>=` negation and `=`/`==` row pairs of :189-212.  file_semantics != 0 applies readLP(File)'s
 * rule of stopping at the first blank line (:76-86).  Returns 0 and malloc'ed arrays (release each
 * with lpsolver_free; names are '\n'-joined in variable order), 1 with the reference's LPException
 * message in err, or 2 for a NumberFormatException-class failure.  Host-only. */
int lpsolver_read_lp(const char* text, int file_semantics, int* m, int* n, int* maximize, double** A,
                     double** b, double** c, char** names, char* err, int err_cap);
void lpsolver_free(void* p);

/* BigDecimal.setScale(6, RoundingMode.HALF_UP).toString() of a binary64 value (LPSolver.java:113);
 * host-only, needs no device.  Returns the string length. */
int lpsolver_set_scale6(double v, char* out, int cap);
/* LPSolver.minInB (LPSolver.java:375-386); host-only. */
int lpsolver_min_in_b(const double* b, int m);

#ifdef __cplusplus
}
#endif
#endif /* LPSOLVER_HOST_H */
