// LPInputReader in the native host layer: the reference's text grammar
// (LPInputReader.java:25-31) read without a JVM, so io_files/input.txt-style files can be solved
// through lpsolver_solve.  Host-only code: no CUDA here.
//
// The three regexes of the reference are restated as one greedy scanner:
//   TERM        := \s* [+-]? \s* \d* \.? \d* \*? [a-zA-Z]+ \d*          (tokenPattern, :30)
//   objective   := TERM+ \s* $                                           (objectivePattern, :26)
//   constraint  := TERM+ \s* (=|==|<=|>=) \s* -? \s* \d+ (\.\d+)? \s* $  (constraintPattern, :27-29)
// For every string those regexes accept the decomposition is unique, so the greedy scan yields
// the same tokens; strings they reject raise the same LPException messages.  The equivalence is
// fuzz-tested against the oracle's `re`-based restatement (tests/test_lp_input_reader.py).
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/lpsolver_host.h"

namespace lpsolver {

namespace {

inline bool is_space(char ch) {  // java \s : [ \t\n\x0B\f\r]
  return ch == ' ' || ch == '\t' || ch == '\n' || ch == '\x0B' || ch == '\f' || ch == '\r';
}
inline bool is_digit(char ch) { return ch >= '0' && ch <= '9'; }
inline bool is_alpha(char ch) { return (ch >= 'a' && ch <= 'z') || (ch >= 'A' && ch <= 'Z'); }

struct Term {
  std::string coef;  // group(2) with whitespace removed
  std::string name;  // group(3)
};

// one TERM starting at pos; on success advances pos and returns true
bool scan_term(const std::string& s, size_t& pos, Term& t) {
  size_t p = pos;
  while (p < s.size() && is_space(s[p])) p++;
  std::string coef;
  if (p < s.size() && (s[p] == '+' || s[p] == '-')) coef.push_back(s[p++]);
  while (p < s.size() && is_space(s[p])) p++;
  while (p < s.size() && is_digit(s[p])) coef.push_back(s[p++]);
  if (p < s.size() && s[p] == '.') coef.push_back(s[p++]);
  while (p < s.size() && is_digit(s[p])) coef.push_back(s[p++]);
  if (p < s.size() && s[p] == '*') p++;
  size_t name0 = p;
  while (p < s.size() && is_alpha(s[p])) p++;
  if (p == name0) return false;
  while (p < s.size() && is_digit(s[p])) p++;
  t.coef = coef;
  t.name = s.substr(name0, p - name0);
  pos = p;
  return true;
}

double parse_decimal(std::string t) {  // new BigDecimal(t) -> binary64 (LPInputReader.java:146-151,181-186)
  if (t.empty() || t == "+") t = "1";
  else if (t == "-") t = "-1";
  bool has_digit = false;
  for (char ch : t) has_digit |= is_digit(ch);
  if (!has_digit) throw std::invalid_argument("NumberFormatException: " + t);  // e.g. "." or "-."
  return std::strtod(t.c_str(), nullptr);
}

}  // namespace

struct ParsedLP {
  LPStandardForm form;
  std::vector<std::string> names;  // variables map: index -> name
};

class LPInputReader {
 public:
  // readLP(String), LPInputReader.java:96-114
  ParsedLP readLP(const std::string& lp) {
    reload();
    std::vector<std::string> lines;
    size_t start = 0;
    for (;;) {
      size_t nl = lp.find('\n', start);
      if (nl == std::string::npos) { lines.push_back(lp.substr(start)); break; }
      lines.push_back(lp.substr(start, nl - start));
      start = nl + 1;
    }
    while (!lines.empty() && lines.back().empty()) lines.pop_back();  // String.split drops trailing empties
    if (lines.size() < 3) throw LPException("Incomplete lp");
    bool maximize = processMaxMinParam(lines[0]);
    processObjective(lines[1]);
    for (size_t i = 2; i < lines.size(); i++) processConstraint(lines[i]);
    return finish(maximize);
  }

  // readLP(File) applied to the file's text, LPInputReader.java:52-93
  ParsedLP readLPFileText(const std::string& text) {
    reload();
    std::vector<std::string> lines;
    size_t start = 0;
    while (start <= text.size()) {
      size_t nl = text.find('\n', start);
      std::string line = text.substr(start, nl == std::string::npos ? std::string::npos : nl - start);
      if (!line.empty() && line.back() == '\r') line.pop_back();
      if (nl == std::string::npos) { if (!line.empty() || start < text.size()) lines.push_back(line); break; }
      lines.push_back(line);
      start = nl + 1;
    }
    if (lines.empty()) throw LPException("Input file is empty");
    bool maximize = processMaxMinParam(lines[0]);
    if (lines.size() < 2) throw std::invalid_argument("objective line missing");
    processObjective(lines[1]);
    int count = 0;
    for (size_t i = 2; i < lines.size(); i++) {
      bool blank = true;
      for (char ch : lines[i]) blank &= is_space(ch);
      if (!blank) { processConstraint(lines[i]); count++; }
      else if (count > 0) break;
      else throw LPException("No constraints in the input file");
    }
    return finish(maximize);
  }

 private:
  std::vector<std::vector<double>> A_;
  std::vector<double> b_, c_;
  std::vector<std::string> names_;
  std::map<std::string, int> index_;

  void reload() { A_.clear(); b_.clear(); c_.clear(); names_.clear(); index_.clear(); }

  static bool processMaxMinParam(const std::string& s) {  // :117-128
    size_t a = 0, b = s.size();
    while (a < b && (unsigned char)s[a] <= ' ') a++;      // String.trim()
    while (b > a && (unsigned char)s[b - 1] <= ' ') b--;
    std::string t = s.substr(a, b - a);
    for (char& ch : t) ch = (char)std::tolower((unsigned char)ch);
    if (t == "min") return false;
    if (t == "max") return true;
    throw LPException("Incorrect max/min parameter");
  }

  void processObjective(const std::string& objective) {    // :131-155
    std::vector<Term> terms;
    size_t pos = 0;
    Term t;
    while (scan_term(objective, pos, t)) terms.push_back(t);
    size_t p = pos;
    while (p < objective.size() && is_space(objective[p])) p++;
    if (terms.empty() || p != objective.size()) throw LPException("Can't recognize objective");
    // variables.put(i, name): a repeated name keeps its LAST index (coefficients.put overwrites)
    std::vector<std::string> var_at(terms.size());
    for (size_t i = 0; i < terms.size(); i++) {
      var_at[i] = terms[i].name;
      index_[terms[i].name] = (int)i;
      c_.push_back(parse_decimal(terms[i].coef));
    }
    names_ = var_at;
  }

  void processConstraint(const std::string& constraint) {  // :158-213
    std::vector<Term> terms;
    size_t pos = 0;
    Term t;
    while (scan_term(constraint, pos, t)) terms.push_back(t);
    size_t p = pos;
    auto skip = [&]() { while (p < constraint.size() && is_space(constraint[p])) p++; };
    bool ok = !terms.empty();
    skip();
    std::string sign;
    if (ok) {
      if (constraint.compare(p, 2, "==") == 0 || constraint.compare(p, 2, "<=") == 0 ||
          constraint.compare(p, 2, ">=") == 0) { sign = constraint.substr(p, 2); p += 2; }
      else if (p < constraint.size() && constraint[p] == '=') { sign = "="; p += 1; }
      else ok = false;
    }
    std::string rhs;
    if (ok) {
      skip();
      if (p < constraint.size() && constraint[p] == '-') rhs.push_back(constraint[p++]);
      skip();
      size_t d0 = p;
      while (p < constraint.size() && is_digit(constraint[p])) rhs.push_back(constraint[p++]);
      if (p == d0) ok = false;
      if (ok && p + 1 < constraint.size() && constraint[p] == '.' && is_digit(constraint[p + 1])) {
        rhs.push_back(constraint[p++]);
        while (p < constraint.size() && is_digit(constraint[p])) rhs.push_back(constraint[p++]);
      }
      skip();
      if (p != constraint.size()) ok = false;
    }
    if (!ok) throw LPException("Can't recognize constraint");
    std::vector<double> row(names_.size(), 0.0);
    for (const Term& tm : terms) {
      auto it = index_.find(tm.name);
      if (it == index_.end()) {                            // a variable first seen in a constraint
        index_[tm.name] = (int)names_.size();
        names_.push_back(tm.name);
        row.push_back(0.0);
        c_.push_back(0.0);
        it = index_.find(tm.name);
      }
      row[(size_t)it->second] = parse_decimal(tm.coef);
    }
    double bnum = std::strtod(rhs.c_str(), nullptr);
    auto negated = [](std::vector<double> r) { for (double& x : r) x = (x == 0.0) ? 0.0 : -x; return r; };
    double nb = (bnum == 0.0) ? 0.0 : -bnum;
    if (sign == ">=") { A_.push_back(negated(row)); b_.push_back(nb); }
    else if (sign == "==" || sign == "=") { A_.push_back(row); A_.push_back(negated(row)); b_.push_back(bnum); b_.push_back(nb); }
    else { A_.push_back(row); b_.push_back(bnum); }
  }

  ParsedLP finish(bool maximize) {                         // normalizeConstraintMatrix :215-223
    ParsedLP out;
    const int n = (int)names_.size(), m = (int)A_.size();
    out.form.m = m;
    out.form.n = n;
    out.form.maximize = maximize;
    std::vector<double> flat((size_t)m * (size_t)n, 0.0);
    for (int i = 0; i < m; i++)
      for (size_t j = 0; j < A_[(size_t)i].size(); j++) flat[(size_t)i * n + j] = A_[(size_t)i][j];
    out.form.setA(std::move(flat));
    out.form.b = b_;
    out.form.c = c_;
    out.form.c.resize((size_t)n, 0.0);
    out.names = names_;
    return out;
  }
};

}  // namespace lpsolver

// ---- flat C wrapper ------------------------------------------------------------------------
extern "C" {

// Parses `text` (file_semantics != 0: readLP(File) rules, stop at the first blank line; else
// readLP(String)).  On success returns 0 and malloc'ed arrays the caller releases with
// lpsolver_free; names are '\n'-joined in variable-index order.  On a grammar error returns 1 and
// puts the reference's exception message into err; 2 = NumberFormatException / other.
int lpsolver_read_lp(const char* text, int file_semantics, int* m, int* n, int* maximize, double** A,
                     double** b, double** c, char** names, char* err, int err_cap) {
  using namespace lpsolver;
  auto set_err = [&](const char* s) { if (err && err_cap > 0) std::snprintf(err, (size_t)err_cap, "%s", s); };
  if (!text || !m || !n || !maximize || !A || !b || !c) { set_err("null argument"); return 2; }
  try {
    LPInputReader reader;
    ParsedLP p = file_semantics ? reader.readLPFileText(text) : reader.readLP(text);
    *m = p.form.m;
    *n = p.form.n;
    *maximize = p.form.maximize ? 1 : 0;
    size_t na = (size_t)p.form.m * (size_t)p.form.n;
    *A = (double*)std::malloc(sizeof(double) * (na ? na : 1));
    *b = (double*)std::malloc(sizeof(double) * (size_t)(p.form.m ? p.form.m : 1));
    *c = (double*)std::malloc(sizeof(double) * (size_t)(p.form.n ? p.form.n : 1));
    if (na) std::memcpy(*A, p.form.A, sizeof(double) * na);
    if (p.form.m) std::memcpy(*b, p.form.b.data(), sizeof(double) * (size_t)p.form.m);
    if (p.form.n) std::memcpy(*c, p.form.c.data(), sizeof(double) * (size_t)p.form.n);
    if (names) {
      std::string joined;
      for (size_t i = 0; i < p.names.size(); i++) { if (i) joined.push_back('\n'); joined += p.names[i]; }
      *names = (char*)std::malloc(joined.size() + 1);
      std::memcpy(*names, joined.c_str(), joined.size() + 1);
    }
    return 0;
  } catch (const LPException& ex) {
    set_err(ex.what());
    return 1;
  } catch (const std::exception& ex) {
    set_err(ex.what());
    return 2;
  }
}

void lpsolver_free(void* p) { std::free(p); }

}  // extern "C"
