"""`LPSolver` (LPSolver.java:15-401) — `LPSolver(...).solve(LPStandardForm)` as a drop-in.

`solve` calls the native host driver (csrc/lp_solver_host.cpp, `lpsolver_solve`), which
sequences phase 1 / phase 2 exactly as the Java class does while every tableau operation
runs on the GPU.  Verdicts come back as the reference's exception types with the reference's
messages.
"""
from __future__ import annotations

import ctypes
import decimal
from ctypes import byref, c_int64, c_void_p
from decimal import Decimal
from typing import List, Optional, Tuple

import numpy as np

from . import _native as N
from .exceptions import LPException, LpsError, SolutionException
from .lp_standard_form import LPStandardForm
from .lp_state import LPState, _dp, _ip


class SolveInfo:
    """What the last solve did (not in the reference API; for parity checks and reporting)."""

    def __init__(self):
        self.verdict = None
        self.used_phase1 = False
        self.x0_index = -1
        self.phase1_log: List[Tuple[int, int]] = []
        self.phase2_log: List[Tuple[int, int]] = []
        self.raw_value: Optional[float] = None
        self.primal: Optional[np.ndarray] = None
        self.device_ms = 0.0
        self.final_state: Optional[LPState] = None


class LPSolver:
    def __init__(self, print_rounder=None, rounder=None, epsilon=LPState.DEF_EPSILON, inf=LPState.DEF_INF,
                 fix_restore_index: bool = False, max_pivots: int = -1, device: int = -1,
                 keep_state: bool = False, log_capacity: int = 1 << 20, loop_mode: int = 0,
                 block_pivots: int = 0):
        # LPSolver.java:24-58.  print_rounder / rounder (java.math.MathContext) are accepted for
        # signature compatibility; device arithmetic is IEEE binary64 round-to-nearest.
        self.print_rounder, self.rounder = print_rounder, rounder
        self.epsilon, self.inf = float(epsilon), float(inf)
        self.fix_restore_index = fix_restore_index
        self.max_pivots = max_pivots
        self.device = device
        self.keep_state = keep_state
        self.log_capacity = log_capacity
        self.loop_mode, self.block_pivots = int(loop_mode), int(block_pivots)   # lps_options, tuning only
        self.info = SolveInfo()

    @staticmethod
    def min_in_b(b) -> int:
        """LPSolver.minInB (LPSolver.java:375-386)."""
        cur, idx = 1e50, -1
        for i, x in enumerate(b):
            if cur > float(x):
                cur, idx = float(x), i
        return idx

    minInB = min_in_b

    def solve(self, form: LPStandardForm) -> Decimal:
        """LPSolver.solve (LPSolver.java:78-94): optimal objective with `setScale(6, HALF_UP)`
        applied (:113), or LPException / SolutionException.  For `min`, `form.c` is negated in
        place as the reference does (:86-89)."""
        lib = N.load()
        opts = N.default_options()
        opts.epsilon, opts.inf, opts.device = self.epsilon, self.inf, self.device
        opts.loop_mode, opts.block_pivots = self.loop_mode, self.block_pivots
        m, n = form.m, form.n
        A = np.ascontiguousarray(form.A, dtype=np.float64).reshape(m, n) if m * n else np.zeros((max(m, 1), max(n, 1)))
        b = np.ascontiguousarray(form.b, dtype=np.float64)
        if form.c.dtype != np.float64 or not form.c.flags["C_CONTIGUOUS"]:
            form.c = np.ascontiguousarray(form.c, dtype=np.float64)
        res = N.LpsolverResult()
        primal = np.zeros(max(n, 1), dtype=np.float64)
        cap = self.log_capacity
        log1 = np.zeros((cap, 2), dtype=np.int32)
        log2 = np.zeros((cap, 2), dtype=np.int32)
        keep = c_void_p()
        rc = lib.lpsolver_solve(byref(opts), m, n, _dp(A), max(n, 1), _dp(b), _dp(form.c), int(form.maximize),
                                int(self.fix_restore_index), int(self.max_pivots), byref(res), _dp(primal),
                                _ip(log1), cap, _ip(log2), cap, byref(keep) if self.keep_state else None)
        info = self.info = SolveInfo()
        info.verdict = res.verdict
        info.used_phase1 = bool(res.used_phase1)
        info.x0_index = res.x0_index
        info.phase1_log = [(int(e), int(l)) for e, l in log1[:min(res.phase1_pivots, cap)]]
        info.phase2_log = [(int(e), int(l)) for e, l in log2[:min(res.phase2_pivots, cap)]]
        info.device_ms = res.device_ms
        if self.keep_state and keep:
            info.final_state = LPState(None, None, None, 0, 0, _handle=keep)
        msg = res.message.decode()
        if res.verdict in (N.LPSOLVER_OPTIMAL, N.LPSOLVER_PIVOT_CAP):
            info.raw_value = res.value
            info.primal = primal[:n].copy()
            return Decimal(res.value6.decode())
        if res.verdict in (N.LPSOLVER_UNBOUNDED, N.LPSOLVER_AUX_UNBOUNDED, N.LPSOLVER_DEGENERATE_FAIL):
            raise SolutionException(msg)
        if res.verdict == N.LPSOLVER_INFEASIBLE:
            raise LPException(msg)
        if res.verdict == N.LPSOLVER_INDEX_ERROR:
            raise IndexError(msg)      # java.lang.ArrayIndexOutOfBoundsException, LPSolver.java:231
        raise LpsError(rc, msg or "lpsolver_solve failed")

    # the reference exposes these two as public (LPSolver.java:248, :283)
    def convert_into_slack_form(self, form: LPStandardForm) -> LPState:
        variables = coefficients = None
        if form.has_variable_names():
            variables, coefficients = form.variables, form.coefficients
            added, counter = 0, 1
            while added < form.m:                       # LPSolver.java:257-266
                name = "x%d" % counter
                if name not in coefficients:
                    variables[form.n + added] = name
                    coefficients[name] = form.n + added
                    added += 1
                counter += 1
        return LPState(form.A, form.b, form.c, form.m, form.n, variables=variables, coefficients=coefficients,
                       epsilon=self.epsilon, inf=self.inf, device=self.device)

    def convert_into_aux_lp(self, form: LPStandardForm) -> LPState:
        st = LPState.aux(form.A, form.b, form.m, form.n, epsilon=self.epsilon, inf=self.inf, device=self.device)
        if form.has_variable_names():
            names = [form.variables[i] for i in range(form.n)]
            x0 = self.get_name_for_x0(form.coefficients)
            names.append(x0)
            used = set(names)
            counter = 1
            while len(names) < form.n + 1 + form.m:
                nm = "x%d" % counter
                if nm not in used:
                    names.append(nm)
                    used.add(nm)
                counter += 1
            st._names0 = names
        return st

    convertIntoSlackForm, convertIntoAuxLP = convert_into_slack_form, convert_into_aux_lp

    @staticmethod
    def get_name_for_x0(coefficients) -> str:
        """LPSolver.getNameForX0 (LPSolver.java:323-342)."""
        if "x0" not in coefficients:
            return "x0"
        if "auxVar" not in coefficients:
            return "auxVar"
        i = 1
        while "auxVar%d" % i in coefficients:
            i += 1
        return "auxVar%d" % i
