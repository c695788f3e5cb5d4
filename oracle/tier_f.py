"""ctypes front end of the C Tier-F oracle (`tier_f.c`) plus a numpy-level driver that restates
LPSolver's phase-1/phase-2 flow on top of it (TEST INFRASTRUCTURE — see oracle/__init__.py).

The driver follows LPSolver.java:78-133 (solve / simplex / initializeSimplex), :135-164
(solveAuxLP), :166-198 (handleInitialization, performDegeneratePivot), :200-246
(restoreInitialLP) and :283-321 (convertIntoAuxLP) with positions tracked in an integer
permutation `pos2var` instead of the two name maps (variable ids: 0..n-1 the structural
variables in input order, then x0 (id n) in the auxiliary LP, then the slacks).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OPTIMAL, UNBOUNDED, PIVOT_CAP = 0, 1, 2


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libtier_f.so")
    src = os.path.join(_HERE, "tier_f.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libtier_f.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        L.tf_get_entering.argtypes = [dp, ctypes.c_int, ctypes.c_double]
        L.tf_get_entering.restype = ctypes.c_int
        L.tf_get_leaving.argtypes = [dp, ctypes.c_long, dp, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_double, ctypes.c_double]
        L.tf_get_leaving.restype = ctypes.c_int
        L.tf_min_in_b.argtypes = [dp, ctypes.c_int, ctypes.c_double]
        L.tf_min_in_b.restype = ctypes.c_int
        L.tf_pivot_tracked.argtypes = [dp, ctypes.c_long, dp, dp, dp, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ip]
        L.tf_pivot_tracked.restype = None
        L.tf_run.argtypes = [dp, ctypes.c_long, dp, dp, dp, ctypes.c_int, ctypes.c_int,
                             ctypes.c_double, ctypes.c_double, ctypes.c_long, ip, ctypes.c_long,
                             ctypes.POINTER(ctypes.c_long), ip, ctypes.c_int]
        L.tf_run.restype = ctypes.c_int
        L.tf_u.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
        L.tf_u.restype = ctypes.c_double
        L.tf_fill_u.argtypes = [dp, ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_long,
                                ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int]
        L.tf_fill_u.restype = None
        L.tf_max_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _dp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a: Optional[np.ndarray]):
    if a is None:
        return ctypes.POINTER(ctypes.c_int)()
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


class TierFState:
    """LPState in binary64 over numpy storage (A row-major C-contiguous, b, c, v)."""

    def __init__(self, A, b, c, v=0.0, eps=1e-9, inf=1e50, nthreads=1, pos2var=None):
        self.A = np.ascontiguousarray(A, dtype=np.float64)
        self.m, self.n = self.A.shape if self.A.ndim == 2 else (0, len(c))
        self.b = np.ascontiguousarray(b, dtype=np.float64)
        self.c = np.ascontiguousarray(c, dtype=np.float64)
        self.v = np.array([v], dtype=np.float64)
        self.eps, self.inf, self.nthreads = eps, inf, nthreads
        if pos2var is None:
            pos2var = np.arange(self.n + self.m, dtype=np.int32)
        self.pos2var = np.ascontiguousarray(pos2var, dtype=np.int32)
        self.log: List[Tuple[int, int]] = []

    @property
    def lda(self):
        return self.A.strides[0] // 8 if self.A.ndim == 2 and self.m > 0 else max(self.n, 1)

    def get_entering(self) -> int:
        return lib().tf_get_entering(_dp(self.c), self.n, self.eps)

    def get_leaving(self, e: int) -> int:
        if not (0 <= e < self.n):
            raise ValueError("entering out of range")
        return lib().tf_get_leaving(_dp(self.A), self.lda, _dp(self.b), self.m, e, self.eps, self.inf)

    def pivot(self, e: int, l: int) -> None:
        lib().tf_pivot_tracked(_dp(self.A), self.lda, _dp(self.b), _dp(self.c), _dp(self.v), self.m,
                               self.n, e, l, self.nthreads, _ip(self.pos2var))
        self.log.append((e, l))

    def run(self, max_pivots: int = -1, log_cap: int = 1 << 22):
        cap = log_cap if max_pivots < 0 else min(log_cap, max_pivots)
        logbuf = np.zeros((max(cap, 1), 2), dtype=np.int32)
        npiv = ctypes.c_long(0)
        status = lib().tf_run(_dp(self.A), self.lda, _dp(self.b), _dp(self.c), _dp(self.v), self.m,
                              self.n, self.eps, self.inf, max_pivots, _ip(logbuf), cap,
                              ctypes.byref(npiv), _ip(self.pos2var), self.nthreads)
        k = npiv.value
        self.log.extend((int(e), int(l)) for e, l in logbuf[:min(k, cap)])
        return status, k

    def position_of(self, var: int) -> int:
        return int(np.nonzero(self.pos2var == var)[0][0])


@dataclass
class TierFResult:
    verdict: str                      # optimal | unbounded | infeasible | pivot_cap | index_error | aux_unbounded | degenerate_fail
    value: Optional[float] = None     # raw v (negated for min), before the 6-decimal rounding
    phase1_log: List[Tuple[int, int]] = field(default_factory=list)
    phase2_log: List[Tuple[int, int]] = field(default_factory=list)
    x0_index: Optional[int] = None
    primal: Optional[np.ndarray] = None
    state: Optional[TierFState] = None
    message: str = ""


def min_in_b(b: np.ndarray, inf: float = 1e50) -> int:
    b = np.ascontiguousarray(b, dtype=np.float64)
    return lib().tf_min_in_b(_dp(b), len(b), inf)


def solve(A, b, c, maximize=True, eps=1e-9, inf=1e50, nthreads=1, max_pivots=-1,
          fix_restore_index=False) -> TierFResult:
    """LPSolver.solve (LPSolver.java:78-94) in binary64.  Inputs are copied."""
    A = np.array(A, dtype=np.float64, order="C")
    b = np.array(b, dtype=np.float64)
    c = np.array(c, dtype=np.float64)
    m, n = len(b), len(c)
    A = A.reshape(m, n)
    if not maximize:
        c = -c
    res = TierFResult("optimal")
    k = min_in_b(b, 1e50)                                   # LPSolver.java:118 uses DEF_INF
    if k == -1 or b[k] >= 0:
        st = TierFState(A, b, c, 0.0, eps, inf, nthreads)
    else:
        # convertIntoAuxLP, LPSolver.java:283-321
        auxA = np.empty((m, n + 1), dtype=np.float64)
        auxA[:, :n] = A
        auxA[:, n] = -1.0
        auxc = np.zeros(n + 1)
        auxc[n] = -1.0
        aux = TierFState(auxA, b.copy(), auxc, 0.0, eps, inf, nthreads)
        x0_var = n
        aux.pivot(n, k)                                     # :138
        status, _ = aux.run(-1)                             # :141-161
        res.phase1_log = list(aux.log)
        if status == UNBOUNDED:
            res.verdict, res.message = "aux_unbounded", "Auxiliary lp is unbounded"
            return res
        x0 = aux.position_of(x0_var)
        res.x0_index = x0
        x0_value = 0.0 if x0 < aux.n else aux.b[x0 - aux.n]  # :169-174
        if abs(x0_value) > eps:
            res.verdict, res.message = "infeasible", "This linear program is infeasible"
            return res
        if x0 >= aux.n:                                     # :182-198
            row = aux.A[x0 - aux.n]
            cand = np.nonzero(np.abs(row) > eps)[0]
            if len(cand) == 0:
                res.verdict, res.message = "degenerate_fail", "Can't perform degenerate pivot"
                return res
            e = int(cand[0])
            aux.pivot(e, x0 - aux.n)
            res.phase1_log = list(aux.log)
            x0 = e
        # restoreInitialLP, LPSolver.java:200-246 (summation in variable-index order)
        newA = np.ascontiguousarray(np.delete(aux.A, x0, axis=1))
        v = 0.0
        newc = np.zeros(n)
        var2pos = np.empty(aux.n + aux.m, dtype=np.int64)
        var2pos[aux.pos2var] = np.arange(aux.n + aux.m)
        for var in range(n):
            coef0 = c[var]
            cur = int(var2pos[var])
            if cur >= aux.n:
                v = v + aux.b[cur - aux.n] * coef0
                newc = newc + (-newA[cur - aux.n]) * coef0
            else:
                kk = cur - 1 if (fix_restore_index and cur > x0) else cur
                if kk >= n:
                    res.verdict, res.message = "index_error", "ArrayIndexOutOfBoundsException: %d" % kk
                    return res
                newc[kk] = newc[kk] + coef0
        pos2var = np.delete(aux.pos2var, x0)
        st = TierFState(newA, aux.b, newc, v, eps, inf, nthreads, pos2var=pos2var)
    status, _ = st.run(max_pivots)
    res.phase2_log = list(st.log)
    res.state = st
    if status == UNBOUNDED:
        res.verdict, res.message = "unbounded", "This linear program is unbounded"
        return res
    if status == PIVOT_CAP:
        res.verdict = "pivot_cap"
    vv = float(st.v[0])
    res.value = vv if maximize else -vv
    var2pos = np.empty(st.n + st.m, dtype=np.int64)
    ids = st.pos2var.copy()
    # after a phase 1 the ids skip x0 (= n); compact them for the lookup
    lookup = {int(v_): p for p, v_ in enumerate(ids)}
    x = np.zeros(n)
    for var in range(n):
        pos = lookup[var]
        if pos >= st.n:
            x[var] = st.b[pos - st.n]
    res.primal = x
    return res


# ---- synthetic inputs -----------------------------------------------------------
_GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x + _GOLDEN
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def u(seed: int, k) -> np.ndarray:
    """u(seed,k) = ((splitmix64(seed ^ (k*GOLDEN)) >> 44) + 1) / 2^20, in (0,1], dyadic."""
    k = np.asarray(k, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = _splitmix64(np.uint64(seed) ^ (k * _GOLDEN))
    return ((h >> np.uint64(44)) + np.uint64(1)).astype(np.float64) / 1048576.0


def gen_dense_feasible(m: int, n: int, seed: int = 0, pos_permille: int = 1000, nthreads: int = 0):
    """C2/C4 family: max c.x, A x <= b, A_ij = u(i*n+j) > 0 (bounded), b_i = (n/4)(1+u) > 0
    (feasible origin).  c_j = +u(mn+j) when (splitmix64(seed ^ ~j) % 1000) < pos_permille else
    -u(mn+j) — pos_permille < 1000 makes the first-positive rule terminate in few pivots."""
    A = np.empty((m, n), dtype=np.float64)
    nt = nthreads or lib().tf_max_threads()
    lib().tf_fill_u(_dp(A), n, 0, m, n, seed, 0, nt)
    j = np.arange(n, dtype=np.uint64)
    c = u(seed, np.uint64(m) * np.uint64(n) + j)
    if pos_permille < 1000:
        sel = _splitmix64(np.uint64(seed) ^ ~j) % np.uint64(1000)
        c = np.where(sel < np.uint64(pos_permille), c, -c)
    i = np.arange(m, dtype=np.uint64)
    b = (n / 4.0) * (1.0 + u(seed, np.uint64(m) * np.uint64(n) + np.uint64(n) + i))
    return A, b, c


def gen_unbounded(m: int, n: int, seed: int = 0, col: int = 0):
    """C5 (iii) family: the dense LP with every c_j > 0 and column `col` of A negated — no positive
    entry in that column, so the LP is unbounded and the loop reports it when the column enters
    (immediately for col = 0, after a long run for col = n-1)."""
    A, b, c = gen_dense_feasible(m, n, seed, 1000)
    A[:, col] = -A[:, col]
    return A, b, c


def gen_assignment(m: int, n: int):
    """C5 (ii) family, degenerate: 0/1 incidence matrix of a bipartite graph (L = m//2 left rows, the
    rest right rows; column j joins left row j % L and right row L + (j % L + 7919 (j // L)) % R),
    b = 1, c = 1.  Totally unimodular: every tableau entry stays in {-1,0,1} (exact in binary64 AND in
    the decimal-15 arithmetic of the reference), with many zero-ratio ties."""
    L = m // 2
    R = m - L
    A = np.zeros((m, n), dtype=np.float64)
    j = np.arange(n, dtype=np.int64)
    A[j % L, j] = 1.0
    A[L + ((j % L) + 7919 * (j // L)) % R, j] = 1.0
    return A, np.ones(m), np.ones(n)


def gen_mixed_rows(m: int, n: int, seed: int = 0, with_equalities: bool = True, frac_bits: int = 10):
    """C3 family: max c.x over rows lowered the way LPInputReader lowers them
    (LPInputReader.java:189-212), with a planted interior point x* = u(.) so the LP is feasible
    but the origin is not (b < 0 on the '>=' rows => phase 1 is forced).
    Rows by i mod 10: 0-5 '<=' with b = A_i.x* + s_i; 6-7 '>=' stored negated: -A_i x <= -(A_i.x* - s_i);
    8-9 one '==' constraint stored as the +/- pair (with_equalities=False turns them into '<=' rows).
    b is rounded to `frac_bits` fractional bits so that every input is a short dyadic number
    (exact in binary64 AND in the decimal-15 oracle)."""
    A = np.empty((m, n), dtype=np.float64)
    nt = lib().tf_max_threads()
    lib().tf_fill_u(_dp(A), n, 0, m, n, seed, 0, nt)
    mn = np.uint64(m) * np.uint64(n)
    j = np.arange(n, dtype=np.uint64)
    i = np.arange(m, dtype=np.uint64)
    c = u(seed, mn + j)
    xstar = u(seed, mn + np.uint64(n) + np.uint64(m) + j)
    slack = (n / 8.0) * u(seed, mn + np.uint64(2 * n) + np.uint64(m) + i)
    scale = float(1 << frac_bits)
    ax = A @ xstar
    b = np.empty(m)
    kind = np.arange(m) % 10
    le = kind <= 5
    ge = (kind == 6) | (kind == 7)
    b[le] = np.round((ax[le] + slack[le]) * scale) / scale
    A[ge] *= -1.0
    b[ge] = np.round((-(ax[ge] - slack[ge])) * scale) / scale
    if with_equalities:
        for r in range(8, m - 1, 10):            # rows r, r+1: the pair +A_r x <= b, -A_r x <= -b
            A[r + 1] = -A[r]
            beq = np.round(ax[r] * scale) / scale
            b[r] = beq
            b[r + 1] = -beq
        if m % 10 == 9:                           # a trailing unpaired row 8 stays a '<=' row
            b[m - 1] = np.round((ax[m - 1] + slack[m - 1]) * scale) / scale
    else:
        eq = kind >= 8
        b[eq] = np.round((ax[eq] + slack[eq]) * scale) / scale
    return A, b, c
