// Host side of the drop-in (include/lpsolver_host.h): LPState / LPSolver mirrors over the C ABI.
// Everything numeric happens on the GPU through lps_*; this file only sequences the calls the
// way LPSolver.java does and maps verdicts to the reference's exception types and messages.
#include "../../include/lpsolver_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

static_assert(sizeof(lpsolver_result) == 256, "lpsolver_result layout");

namespace lpsolver {

// ---- LPState -------------------------------------------------------------------------------
void LPState::check(int rc, const char* what) {
  if (rc == LPS_OK) return;
  std::string msg = std::string(what) + ": " + lps_status_string(rc);
  if (h_) msg += std::string(" (") + lps_last_error(h_) + ")";
  if (rc == LPS_ERR_INVALID) throw std::invalid_argument(msg);  // Validate.isTrue, LPState.java:288
  throw SolutionException(msg);
}

LPState::LPState(const double* A, int64_t lda, const double* b, const double* c, double v, int m,
                 int n, const lps_options& opts) {
  int rc = lps_create(&h_, &opts);
  if (rc != LPS_OK) throw SolutionException(std::string("lps_create: ") + lps_status_string(rc));
  check(lps_load(h_, m, n, A, lda, b, c, v), "lps_load");
}

LPState LPState::aux(const double* A, int64_t lda, const double* b, int m, int n,
                     const lps_options& opts) {
  lps_handle h = nullptr;
  int rc = lps_create(&h, &opts);
  if (rc != LPS_OK) throw SolutionException(std::string("lps_create: ") + lps_status_string(rc));
  LPState st(h);
  st.check(lps_load_aux(h, m, n, A, lda, b), "lps_load_aux");
  return st;
}

LPState::~LPState() {
  if (h_) lps_destroy(h_);
}
LPState::LPState(LPState&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }

int LPState::getEntering() {
  int e = -1;
  check(lps_get_entering(h_, &e), "getEntering");
  return e;
}
int LPState::getLeaving(int entering) {
  int l = -1;
  check(lps_get_leaving(h_, entering, &l), "getLeaving");
  return l;
}
void LPState::pivot(int entering, int leaving) { check(lps_pivot(h_, entering, leaving), "pivot"); }
lps_run_result LPState::run(int64_t max_pivots) {
  lps_run_result r;
  check(lps_run(h_, max_pivots, &r), "run");
  return r;
}
int LPState::m() const { int m = 0, n = 0; lps_dims(h_, &m, &n); return m; }
int LPState::n() const { int m = 0, n = 0; lps_dims(h_, &m, &n); return n; }
double LPState::v() { double v = 0; check(lps_read_v(h_, &v), "read v"); return v; }
std::vector<double> LPState::b() {
  std::vector<double> out((size_t)m());
  if (!out.empty()) check(lps_read_b(h_, out.data()), "read b");
  return out;
}
std::vector<double> LPState::c() {
  std::vector<double> out((size_t)n());
  if (!out.empty()) check(lps_read_c(h_, out.data()), "read c");
  return out;
}
std::vector<double> LPState::row(int i) {
  std::vector<double> out((size_t)n());
  check(lps_read_row(h_, i, out.data()), "read row");
  return out;
}
std::vector<int> LPState::positions() {
  std::vector<int> out((size_t)(m() + n()));
  if (!out.empty()) check(lps_read_positions(h_, out.data()), "read positions");
  return out;
}
int LPState::positionOf(int var) {
  int pos = -1;
  check(lps_position_of(h_, var, &pos), "position_of");
  return pos;
}
std::vector<std::pair<int, int>> LPState::pivotLog() {
  int64_t cnt = 0;
  check(lps_read_pivot_log(h_, nullptr, 0, &cnt), "pivot log");
  std::vector<int> raw((size_t)(2 * cnt));
  if (cnt) check(lps_read_pivot_log(h_, raw.data(), cnt, &cnt), "pivot log");
  std::vector<std::pair<int, int>> out((size_t)cnt);
  for (int64_t k = 0; k < cnt; k++) out[(size_t)k] = {raw[2 * k], raw[2 * k + 1]};
  return out;
}

// ---- LPSolver ------------------------------------------------------------------------------
LPSolver::LPSolver() { lps_default_options(&options); }
LPSolver::LPSolver(double epsilon, double inf) {
  lps_default_options(&options);
  options.epsilon = epsilon;
  options.inf = inf;
}

int LPSolver::minInB(const std::vector<double>& b) {
  // LPSolver.java:375-386: first index of the strict minimum, start value DEF_INF
  double cur = 1e50;
  int idx = -1;
  for (size_t i = 0; i < b.size(); i++)
    if (cur > b[i]) {
      cur = b[i];
      idx = (int)i;
    }
  return idx;
}

std::string LPSolver::setScale6(double v) {
  // BigDecimal.setScale(6, HALF_UP) of the exact binary64 value (LPSolver.java:113).  glibc
  // prints the exact decimal expansion; HALF_UP only needs the 7th fractional digit.
  if (!std::isfinite(v)) return std::isnan(v) ? "NaN" : (v > 0 ? "Infinity" : "-Infinity");
  std::vector<char> buf(1600);
  std::snprintf(buf.data(), buf.size(), "%.1100f", std::fabs(v));
  std::string s(buf.data());
  size_t dot = s.find('.');
  std::string digits = s.substr(0, dot) + s.substr(dot + 1, 6);  // integer part + 6 decimals
  bool up = s[dot + 7] >= '5';
  if (up) {
    int k = (int)digits.size() - 1;
    while (k >= 0 && digits[(size_t)k] == '9') digits[(size_t)k--] = '0';
    if (k >= 0) digits[(size_t)k]++; else digits.insert(digits.begin(), '1');
  }
  std::string out = digits.substr(0, digits.size() - 6) + "." + digits.substr(digits.size() - 6);
  bool nonzero = out.find_first_not_of("0.") != std::string::npos;
  if (v < 0 && nonzero) out.insert(out.begin(), '-');
  return out;
}

LPState LPSolver::initializeSimplex(LPStandardForm& f) {
  // LPSolver.java:116-133
  const int m = f.m, n = f.n;
  int k = minInB(f.b);
  if (k == -1 || f.b[(size_t)k] >= 0.0) {
    // convertIntoSlackForm, LPSolver.java:248-272
    return LPState(f.A, f.lda, f.b.data(), f.c.data(), 0.0, m, n, options);
  }
  trace.used_phase1 = true;
  // convertIntoAuxLP (:283-321) happens on the device while loading
  LPState aux = LPState::aux(f.A, f.lda, f.b.data(), m, n, options);
  const int n_aux = n + 1;
  const int x0_var = n;
  // solveAuxLP, :135-164: forced first pivot, then the loop (on the device)
  aux.pivot(n_aux - 1, k);
  lps_run_result r = aux.run(-1);
  trace.phase1_pivots = r.total_pivots;
  trace.device_ms += r.device_ms;
  if (r.verdict == LPS_UNBOUNDED) {
    trace.phase1_log = aux.pivotLog();
    throw SolutionException("Auxiliary lp is unbounded");  // :149
  }
  int x0 = aux.positionOf(x0_var);                          // the reference tracks it pivot by pivot (:151-155)
  trace.x0_index = x0;
  // handleInitialization, :166-180
  std::vector<double> bb = aux.b();
  double x0_value = (x0 < n_aux) ? 0.0 : bb[(size_t)(x0 - n_aux)];
  if (std::fabs(x0_value) > options.epsilon) {
    trace.phase1_log = aux.pivotLog();
    throw LPException("This linear program is infeasible");  // :173
  }
  if (x0 >= n_aux) {
    // performDegeneratePivot, :182-198
    int row = x0 - n_aux, e = -1;
    int rc = lps_first_nonzero_in_row(aux.handle(), row, &e);
    if (rc != LPS_OK) throw SolutionException(lps_last_error(aux.handle()));
    if (e == -1) throw SolutionException("Can't perform degenerate pivot");  // :193
    aux.pivot(e, row);
    trace.phase1_pivots += 1;
    x0 = e;
  }
  trace.phase1_log = aux.pivotLog();
  // restoreInitialLP, :200-246.  Summation order: ascending variable index (the reference
  // iterates a HashMap's keySet).
  std::vector<int> pos = aux.positions();
  std::vector<int> var2pos((size_t)(n_aux + m), -1);
  for (size_t p = 0; p < pos.size(); p++) var2pos[(size_t)pos[p]] = (int)p;
  std::vector<lps_objective_op> ops;
  ops.reserve((size_t)n);
  for (int var = 0; var < n; var++) {
    lps_objective_op op;
    op.coef = f.c[(size_t)var];
    int cur = var2pos[(size_t)var];
    if (cur >= n_aux) {
      op.kind = 0;
      op.index = cur - n_aux;
    } else {
      int kk = (fix_restore_index && cur > x0) ? cur - 1 : cur;  // :231 uses `cur` unshifted
      if (kk >= n) throw ArrayIndexOutOfBounds("ArrayIndexOutOfBoundsException: " + std::to_string(kk));
      op.kind = 1;
      op.index = kk;
    }
    ops.push_back(op);
  }
  int rc = lps_drop_column(aux.handle(), x0);
  if (rc == LPS_OK) rc = lps_rebuild_objective(aux.handle(), ops.data(), (int)ops.size());
  if (rc != LPS_OK) throw SolutionException(lps_last_error(aux.handle()));
  return aux;
}

double LPSolver::solve(LPStandardForm& f) {
  // LPSolver.java:78-94 (+ simplex :96-114)
  trace = SolveTrace();
  primal.clear();
  if (!f.maximize)
    for (double& x : f.c) x = -x;  // negated IN PLACE, :86-89
  LPState st = initializeSimplex(f);
  const int64_t before = trace.phase1_pivots;
  lps_run_result r = st.run(max_pivots);
  trace.device_ms += r.device_ms;
  trace.phase2_pivots = r.npivots;
  {
    auto log = st.pivotLog();
    size_t skip = std::min<size_t>((size_t)before, log.size());
    trace.phase2_log.assign(log.begin() + (trace.used_phase1 ? (long)skip : 0), log.end());
  }
  trace.verdict = r.verdict;
  if (r.verdict == LPS_UNBOUNDED) {
    if (keep) keep_state = st.release();
    throw SolutionException("This linear program is unbounded");  // :105
  }
  trace.raw_v = r.v;
  primal.assign((size_t)f.n, 0.0);
  if (f.n > 0) lps_read_primal(st.handle(), f.n, primal.data());
  if (keep) keep_state = st.release();
  return f.maximize ? r.v : -r.v;
}

}  // namespace lpsolver

// ---- flat C wrapper ------------------------------------------------------------------------
extern "C" int lpsolver_solve(const lps_options* opts, int m, int n, const double* A, int64_t lda,
                              const double* b, double* c, int maximize, int fix_restore_index,
                              int64_t max_pivots, lpsolver_result* res, double* primal,
                              int* phase1_log, int64_t phase1_cap, int* phase2_log,
                              int64_t phase2_cap, lps_handle* keep_state) {
  using namespace lpsolver;
  if (!res || m < 0 || n < 0 || lda < n) return LPS_ERR_INVALID;
  std::memset(res, 0, sizeof(*res));
  res->x0_index = -1;
  LPSolver solver;
  if (opts) solver.options = *opts;
  solver.fix_restore_index = fix_restore_index != 0;
  solver.max_pivots = max_pivots;
  solver.keep = keep_state != nullptr;
  LPStandardForm f;
  f.m = m;
  f.n = n;
  f.maximize = maximize != 0;
  f.A = A;     // a view: the tableau is copied once, host -> HBM, inside lps_load
  f.lda = lda;
  f.b.assign(b, b + m);
  f.c.assign(c, c + n);
  auto set_msg = [&](const char* s) { std::snprintf(res->message, sizeof(res->message), "%s", s); };
  auto copy_logs = [&]() {
    const SolveTrace& t = solver.trace;
    res->used_phase1 = t.used_phase1;
    res->x0_index = t.x0_index;
    res->phase1_pivots = (int64_t)t.phase1_log.size();
    res->phase2_pivots = (int64_t)t.phase2_log.size();
    res->device_ms = t.device_ms;
    if (phase1_log)
      for (int64_t k = 0; k < std::min<int64_t>(phase1_cap, (int64_t)t.phase1_log.size()); k++) {
        phase1_log[2 * k] = t.phase1_log[(size_t)k].first;
        phase1_log[2 * k + 1] = t.phase1_log[(size_t)k].second;
      }
    if (phase2_log)
      for (int64_t k = 0; k < std::min<int64_t>(phase2_cap, (int64_t)t.phase2_log.size()); k++) {
        phase2_log[2 * k] = t.phase2_log[(size_t)k].first;
        phase2_log[2 * k + 1] = t.phase2_log[(size_t)k].second;
      }
    if (keep_state) *keep_state = solver.keep_state;
  };
  int rc = LPS_OK;
  try {
    double val = solver.solve(f);
    res->verdict = solver.trace.verdict == LPS_PIVOT_CAP ? LPSOLVER_PIVOT_CAP : LPSOLVER_OPTIMAL;
    res->value = val;
    std::string s6 = LPSolver::setScale6(val);
    std::snprintf(res->value6, sizeof(res->value6), "%s", s6.c_str());
    if (primal) std::copy(solver.primal.begin(), solver.primal.end(), primal);
  } catch (const SolutionException& ex) {
    std::string msg = ex.what();
    if (msg == "This linear program is unbounded") res->verdict = LPSOLVER_UNBOUNDED;
    else if (msg == "Auxiliary lp is unbounded") res->verdict = LPSOLVER_AUX_UNBOUNDED;
    else if (msg == "Can't perform degenerate pivot") res->verdict = LPSOLVER_DEGENERATE_FAIL;
    else { res->verdict = LPSOLVER_ERROR; rc = LPS_ERR_CUDA; }
    set_msg(ex.what());
  } catch (const LPException& ex) {
    res->verdict = LPSOLVER_INFEASIBLE;
    set_msg(ex.what());
  } catch (const ArrayIndexOutOfBounds& ex) {
    res->verdict = LPSOLVER_INDEX_ERROR;
    set_msg(ex.what());
  } catch (const std::exception& ex) {
    res->verdict = LPSOLVER_ERROR;
    set_msg(ex.what());
    rc = LPS_ERR_INVALID;
  }
  for (int j = 0; j < n; j++) c[j] = f.c[(size_t)j];  // the in-place negation for min (:86-89)
  copy_logs();
  return rc;
}

extern "C" int lpsolver_set_scale6(double v, char* out, int cap) {
  std::string s = lpsolver::LPSolver::setScale6(v);
  if (out && cap > 0) std::snprintf(out, (size_t)cap, "%s", s.c_str());
  return (int)s.size();
}

extern "C" int lpsolver_min_in_b(const double* b, int m) {
  if (m <= 0 || !b) return -1;
  return lpsolver::LPSolver::minInB(std::vector<double>(b, b + m));
}
