"""Restatement of the reference's simplex hot path and its driver, generic over the number
system (TEST INFRASTRUCTURE — see oracle/__init__.py).

Every function cites the reference lines it follows (paths relative to
`src/main/java/lpsolver/` of Toptachamann/Linear_Programming_Solver).  With `arith.Dec15`
this is Tier D (the reference's own arithmetic); with `arith.F64` it is the Python form of
Tier F (the binary64 twin the GPU must match bit for bit; `tier_f.c` is the fast C form).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

from .arith import Dec15, F64  # noqa: F401  (re-exported)
from .java_hashmap import hashmap_key_order


class LPException(Exception):
    """LPException.java:4-14."""


class SolutionException(LPException):
    """SolutionException.java:3-13 (extends LPException)."""


class LPStandardForm:
    """LPStandardForm.java:10-65 — plain holder of A, b, c, m, n, maximize and the two name maps."""

    def __init__(self, A, b, c, m, n, maximize, variables=None, coefficients=None, arith=Dec15):
        self.arith = arith
        cv = arith.conv
        self.A = [[cv(x) for x in row] for row in A]
        self.b = [cv(x) for x in b]
        self.c = [cv(x) for x in c]
        self.m = m
        self.n = n
        self.maximize = maximize
        self.variables: Optional[Dict[int, str]] = variables
        self.coefficients: Optional[Dict[str, int]] = coefficients
        # Iteration order of `coefficients.keySet()` matters in restoreInitialLP
        # (LPSolver.java:213-233).  "java" = java.util.HashMap order, "insertion" = the order
        # of this dict (what a Groovy map literal — a LinkedHashMap — gives in the Spock
        # specs), "index" = ascending variable index (what the GPU build uses).
        self.key_order = "java"

    def has_variable_names(self):  # LPStandardForm.java:154-156
        return self.variables is not None and self.coefficients is not None


class LPState:
    """LPState.java:17-320 — slack-form tableau A[m][n], b[m], c[n], v and the pivot rules."""

    def __init__(self, A, b, c, m, n, v=None, variables=None, coefficients=None, arith=Dec15,
                 epsilon=None, inf=None):
        # LPState.java:88-99: the arrays are aliased, not copied; v defaults to ZERO;
        # rounder/epsilon/INF default to DEF_* — which is what every LPSolver call site gets
        # (LPSolver.java:245,267,270).
        self.arith = arith
        self.A = A
        self.b = b
        self.c = c
        self.v = arith.ZERO if v is None else v
        self.m = m
        self.n = n
        self.variables = variables
        self.coefficients = coefficients
        self.epsilon = arith.DEF_EPSILON if epsilon is None else epsilon
        self.INF = arith.DEF_INF if inf is None else inf
        self.pivot_log: List[Tuple[int, int]] = []
        # what LPState.java:115-118 prints at TRACE level for each pivot: the entering NAME
        # and `variables.get(leaving)` — the name at NON-BASIC position `leaving` (the
        # reference logs the wrong name but it pins the leaving row index) — or the raw
        # indices when the state has no names.  Used to check the recorded runs in
        # logs/lp_solver.log.
        self.name_log: List[Tuple[object, object]] = []

    # -- selection ---------------------------------------------------------------------
    def get_entering(self) -> int:
        """LPState.java:274-285: first index with c[i] > epsilon, else -1."""
        cmp, eps = self.arith.cmp, self.epsilon
        for i in range(self.n):
            if cmp(self.c[i], eps) > 0:
                return i
        return -1

    def get_leaving(self, entering: int) -> int:
        """LPState.java:287-305: lowest-index minimum of b[i]/A[i][e] over rows with
        A[i][e] >= epsilon; strict '<' against a running minimum that starts at INF."""
        if not (0 <= entering < self.n):  # Validate.isTrue, LPState.java:288
            raise ValueError("entering out of range")
        ar = self.arith
        leaving = -1
        min_slack = self.INF
        for i in range(self.m):
            aie = self.A[i][entering]
            if ar.cmp(aie, self.epsilon) < 0:
                slack = self.INF
            else:
                slack = ar.div(self.b[i], aie)
            if ar.cmp(slack, min_slack) < 0:
                min_slack = slack
                leaving = i
        return leaving

    # -- pivot ---------------------------------------------------------------------------
    def pivot(self, entering: int, leaving: int) -> None:
        """LPState.java:114-181 (`pivotSequentially`; `pivotConcurrently` :184-272 computes the
        same cells, each with a single writer, so the values are identical)."""
        ar = self.arith
        mul, sub, add, div, neg = ar.mul, ar.sub, ar.add, ar.div, ar.neg
        A, b, c, n, m = self.A, self.b, self.c, self.n, self.m
        if self.variables is not None and self.coefficients is not None:
            self.name_log.append((self.variables.get(entering), self.variables.get(leaving)))
        else:
            self.name_log.append((entering, leaving))
        pivot_row = A[leaving]
        p = pivot_row[entering]                       # :138
        pivot_row[entering] = div(ar.ONE, p)          # :139
        for j in range(n):                            # :140-145
            if j != entering:
                pivot_row[j] = div(pivot_row[j], p)
        b[leaving] = div(b[leaving], p)               # :146
        b_ent = b[leaving]                            # :150
        for i in range(m):                            # :151-166
            if i == leaving:
                continue
            row = A[i]
            a = row[entering]                         # :156
            row[entering] = neg(div(a, p))            # :157
            for j in range(n):
                if j != entering:
                    row[j] = sub(row[j], mul(a, pivot_row[j]))   # :162
            b[i] = sub(b[i], mul(a, b_ent))           # :164
        ce = c[entering]                              # :170
        self.v = add(self.v, mul(b[leaving], ce))     # :171
        c[entering] = neg(div(ce, p))                 # :172
        for j in range(n):                            # :173-178
            if j != entering:
                c[j] = sub(c[j], mul(ce, pivot_row[j]))
        self._exchange_indexes(entering, leaving)     # :180
        self.pivot_log.append((entering, leaving))

    def _exchange_indexes(self, entering: int, leaving: int) -> None:
        """LPState.java:311-320."""
        if self.variables is not None and self.coefficients is not None:
            n = self.n
            ent_name = self.variables.get(entering)
            leav_name = self.variables.get(leaving + n)
            self.variables[entering] = leav_name
            self.variables[leaving + n] = ent_name
            self.coefficients[ent_name] = leaving + n
            self.coefficients[leav_name] = entering


class SolveTrace:
    """What a solve did, for parity checks (not part of the reference API)."""

    def __init__(self):
        self.phase1_log: List[Tuple[int, int]] = []
        self.phase2_log: List[Tuple[int, int]] = []
        self.aux_state: Optional[LPState] = None
        self.used_phase1 = False
        self.x0_final_index: Optional[int] = None
        self.final_state: Optional[LPState] = None
        self.raw_v = None


class LPSolver:
    """LPSolver.java:15-401 — the driver around LPState."""

    def __init__(self, arith=Dec15, epsilon=None, inf=None, max_pivots: Optional[int] = None,
                 fix_restore_index: bool = False):
        self.arith = arith
        # The reference indexes the rebuilt objective with the variable's position in the
        # AUXILIARY tableau (LPSolver.java:220,231) although column x0 has just been removed
        # (:205-211): a non-basic original variable sitting to the right of x0 lands one
        # column too far right (or throws ArrayIndexOutOfBoundsException at position n).
        # False = restate the reference as written; True = shift the index.
        self.fix_restore_index = fix_restore_index
        # LPSolver.java:24-58: ctor options.  Only `epsilon` (handleInitialization :171,
        # performDegeneratePivot :187) and `rounder` (restoreInitialLP :223-231) are live.
        self.epsilon = arith.DEF_EPSILON if epsilon is None else epsilon
        self.inf = arith.DEF_INF if inf is None else inf
        self.max_pivots = max_pivots  # oracle-only safety cap
        self.trace = SolveTrace()

    # -- public -----------------------------------------------------------------------------
    def solve(self, st: LPStandardForm):
        """LPSolver.java:78-94.  For `min` the objective is negated IN PLACE (:86-89)."""
        self.trace = SolveTrace()
        if st.maximize:
            return self._simplex(st)
        for i in range(len(st.c)):
            st.c[i] = self.arith.neg(st.c[i])
        res = self._simplex(st)
        return -res if res else abs(res)

    # -- driver -------------------------------------------------------------------------
    def _simplex(self, st: LPStandardForm):
        """LPSolver.java:96-114."""
        state = self._initialize_simplex(st)
        k = 0
        while True:
            e = state.get_entering()
            if e == -1:
                break
            l = state.get_leaving(e)
            if l == -1:
                self.trace.phase2_log = list(state.pivot_log)
                self.trace.final_state = state
                raise SolutionException("This linear program is unbounded")  # :105
            state.pivot(e, l)
            k += 1
            if self.max_pivots is not None and k >= self.max_pivots:
                break
        self.trace.phase2_log = list(state.pivot_log)
        self.trace.final_state = state
        self.trace.raw_v = state.v
        return self.arith.set_scale6(state.v)  # :113

    def _initialize_simplex(self, st: LPStandardForm) -> LPState:
        """LPSolver.java:116-133."""
        k = self.min_in_b(st.b)
        if k == -1 or self.arith.cmp(st.b[k], self.arith.ZERO) >= 0:
            return self.convert_into_slack_form(st)
        self.trace.used_phase1 = True
        if not st.has_variable_names():
            self._add_default_variables(st)
        aux = self.convert_into_aux_lp(st)
        self.trace.aux_state = aux
        x0 = self.solve_aux_lp(aux, aux.n - 1, k)
        self.trace.phase1_log = list(aux.pivot_log)
        return self.handle_initialization(aux, st, x0)

    def solve_aux_lp(self, aux: LPState, index_of_x0: int, min_in_b: int) -> int:
        """LPSolver.java:135-164."""
        n = aux.n
        aux.pivot(index_of_x0, min_in_b)        # :138 forced first pivot
        x0 = min_in_b + n                        # :139
        while True:
            e = aux.get_entering()
            if e == -1:
                break
            l = aux.get_leaving(e)
            if l == -1:
                raise SolutionException("Auxiliary lp is unbounded")  # :149
            if e == x0:                          # :151-155
                x0 = l + n
            elif l + n == x0:
                x0 = e
            aux.pivot(e, l)
        self.trace.x0_final_index = x0
        return x0

    def handle_initialization(self, aux: LPState, initial: LPStandardForm, x0: int) -> LPState:
        """LPSolver.java:166-180."""
        ar = self.arith
        x0_value = ar.ZERO if x0 < aux.n else aux.b[x0 - aux.n]
        if ar.cmp(ar.abs(x0_value), self.epsilon) > 0:
            raise LPException("This linear program is infeasible")  # :173
        if x0 >= aux.n:
            x0 = self.perform_degenerate_pivot(aux, x0)
            self.trace.phase1_log = list(aux.pivot_log)
        return self.restore_initial_lp(aux, initial, x0)

    def perform_degenerate_pivot(self, aux: LPState, index_of_x0: int) -> int:
        """LPSolver.java:182-198."""
        ar = self.arith
        row = aux.A[index_of_x0 - aux.n]
        entering = -1
        for i in range(aux.n):
            if ar.cmp(ar.abs(row[i]), self.epsilon) > 0:
                entering = i
                break
        if entering == -1:
            raise SolutionException("Can't perform degenerate pivot")  # :193
        aux.pivot(entering, index_of_x0 - aux.n)
        return entering

    def restore_initial_lp(self, aux: LPState, initial: LPStandardForm, index_of_x0: int) -> LPState:
        """LPSolver.java:200-246: drop column x0, rebuild c and v by substituting the basic
        original variables, shift the name maps down by one."""
        ar = self.arith
        n = initial.n
        m = aux.m
        A = [row[:index_of_x0] + row[index_of_x0 + 1: n + 1] for row in aux.A]   # :205-211
        v = ar.ZERO
        c = [ar.ZERO] * n
        if initial.key_order == "java":
            keys = hashmap_key_order(list(initial.coefficients.keys()))
        elif initial.key_order == "index":
            keys = sorted(initial.coefficients.keys(), key=lambda k: initial.coefficients[k])
        else:
            keys = list(initial.coefficients.keys())
        for var in keys:                                                     # :217-233
            index = initial.coefficients[var]
            coef0 = initial.c[index]
            cur = aux.coefficients[var]
            if cur >= aux.n:
                v = ar.add(v, ar.mul(aux.b[cur - aux.n], coef0))             # :223
                row = A[cur - aux.n]
                for j in range(n):
                    c[j] = ar.add(c[j], ar.mul(ar.neg(row[j]), coef0))        # :226-227
            else:
                k = cur - 1 if (self.fix_restore_index and cur > index_of_x0) else cur
                if k >= n:
                    raise IndexError("ArrayIndexOutOfBoundsException: %d" % k)  # c has length n
                c[k] = ar.add(c[k], coef0)                                   # :231
        variables, coefficients = aux.variables, aux.coefficients            # :235-244
        x0_name = variables.get(index_of_x0)
        coefficients.pop(x0_name, None)
        for i in range(index_of_x0, n + m):
            name = variables.get(i + 1)
            variables[i] = name
            coefficients[name] = i
        variables.pop(n + m, None)
        return LPState(A, aux.b, c, initial.m, initial.n, v=v, variables=variables,
                       coefficients=coefficients, arith=ar)

    def convert_into_slack_form(self, st: LPStandardForm) -> LPState:
        """LPSolver.java:248-272: arrays are ALIASED into the LPState."""
        if st.has_variable_names():
            coefficients, variables = st.coefficients, st.variables
            m, n = st.m, st.n
            added, counter = 0, 1
            while added < m:
                name = "x" + str(counter)
                if name not in coefficients:
                    variables[n + added] = name
                    coefficients[name] = n + added
                    added += 1
                counter += 1
            return LPState(st.A, st.b, st.c, m, n, variables=variables, coefficients=coefficients,
                           arith=self.arith)
        return LPState(st.A, st.b, st.c, st.m, st.n, arith=self.arith)

    def convert_into_aux_lp(self, st: LPStandardForm) -> LPState:
        """LPSolver.java:283-321: copy A into m x (n+1) with a last column of -1, c_aux =
        (0,…,0,-1), copy b and the name maps, name x0."""
        ar = self.arith
        m, n = st.m, st.n
        minus1 = ar.neg(ar.ONE)
        aux_a = [list(st.A[i][:n]) + [minus1] for i in range(m)]
        b = list(st.b[:m])
        aux_c = [ar.ZERO] * n + [minus1]
        variables = dict(st.variables)
        coefficients = dict(st.coefficients)
        x0 = self.get_name_for_x0(st.coefficients)
        variables[n] = x0
        coefficients[x0] = n
        aux_form = LPStandardForm(aux_a, b, aux_c, m, n + 1, st.maximize, variables, coefficients,
                                  arith=ar)
        return self.convert_into_slack_form(aux_form)

    @staticmethod
    def get_name_for_x0(coefficients) -> str:
        """LPSolver.java:323-342."""
        if "x0" not in coefficients:
            return "x0"
        if "auxVar" not in coefficients:
            return "auxVar"
        i = 1
        while "auxVar" + str(i) in coefficients:
            i += 1
        return "auxVar" + str(i)

    def min_in_b(self, b: Sequence) -> int:
        """LPSolver.java:375-386: first index of the strict minimum, start value DEF_INF."""
        ar = self.arith
        cur = ar.DEF_INF
        idx = -1
        for i in range(len(b)):
            if ar.cmp(cur, b[i]) > 0:
                cur = b[i]
                idx = i
        return idx

    @staticmethod
    def _add_default_variables(st: LPStandardForm) -> None:
        """LPSolver.java:388-400."""
        st.variables = {i: "x" + str(i + 1) for i in range(st.n)}
        st.coefficients = {"x" + str(i + 1): i for i in range(st.n)}


def primal_solution(state: LPState, n_original: int, name_of=None) -> List:
    """Primal values of the original variables from a final state: x[var] = b[pos-n] if the
    variable sits at a basic position pos >= n, else 0 (what `io_files/output.txt:214-233`
    prints; the reference's own printer is commented out at LPSolver.java:344-374)."""
    ar = state.arith
    x = [ar.ZERO] * n_original
    if state.coefficients is None:
        raise ValueError("state has no variable names; use pos2var bookkeeping instead")
    for k in range(n_original):
        name = name_of(k) if name_of else "x" + str(k + 1)
        pos = state.coefficients[name]
        if pos >= state.n:
            x[k] = state.b[pos - state.n]
    return x
