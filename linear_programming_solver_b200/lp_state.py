"""`LPState` (LPState.java:17-320) with the tableau resident in HBM.

Same method set as the reference class — `get_entering`, `get_leaving`, `pivot` (camelCase
aliases included) — plus `run`, the reference's pivot loops (LPSolver.java:101-112, :141-161)
executed on the device.  Fields `A, b, c, v, m, n` are read back on access.  Name maps
(`variables`, `coefficients`) are kept on the host and derived from the device's position
permutation, which is what `exchangeIndexes` (LPState.java:311-320) maintains.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_int, c_int64, c_void_p
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _native as N
from .exceptions import LpsError, SolutionException


def _dp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(c_double))


def _ip(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(c_int))


class LPState:
    DEF_EPSILON = 1e-9   # LPState.java:20
    DEF_INF = 1e50       # LPState.java:21

    def __init__(self, A, b, c, m: int, n: int, v: float = 0.0,
                 variables: Optional[Dict[int, str]] = None,
                 coefficients: Optional[Dict[str, int]] = None,
                 epsilon: float = DEF_EPSILON, inf: float = DEF_INF, device: int = -1,
                 time_kernels: bool = False, loop_mode: int = 0, block_pivots: int = 0, _handle=None,
                 _aux=False, update_variant: int = -1, panel_ctas: int = 0, pass_chunk_rows: int = 0):
        self._lib = N.load()
        self._h = c_void_p()
        self._names0 = None
        if variables is not None and coefficients is not None:
            self._names0 = [variables.get(i) for i in range(m + n)]
        if _handle is not None:
            self._h = _handle
            return
        opts = N.default_options()
        opts.epsilon, opts.inf, opts.device, opts.time_kernels = epsilon, inf, device, int(time_kernels)
        opts.loop_mode = int(loop_mode)
        opts.block_pivots = int(block_pivots)
        opts.update_variant = int(update_variant)
        opts.panel_ctas, opts.pass_chunk_rows = int(panel_ctas), int(pass_chunk_rows)   # tuning only
        rc = self._lib.lps_create(byref(self._h), byref(opts))
        if rc != N.LPS_OK:
            raise LpsError(rc, "lps_create: " + self._lib.lps_status_string(rc).decode())
        A = np.ascontiguousarray(np.asarray(A, dtype=np.float64).reshape(m, n)) if m * n else np.zeros((m, max(n, 1)))
        b = np.ascontiguousarray(np.asarray(b, dtype=np.float64).reshape(m))
        if _aux:
            self._ck(self._lib.lps_load_aux(self._h, m, n, _dp(A), A.strides[0] // 8 if m else n, _dp(b)), "lps_load_aux")
        else:
            c = np.ascontiguousarray(np.asarray(c, dtype=np.float64).reshape(n))
            self._ck(self._lib.lps_load(self._h, m, n, _dp(A), A.strides[0] // 8 if m else max(n, 1), _dp(b), _dp(c),
                                        float(v)), "lps_load")

    # -- construction helpers ------------------------------------------------------------
    @classmethod
    def aux(cls, A, b, m, n, **kw) -> "LPState":
        """LPSolver.convertIntoAuxLP (LPSolver.java:283-321), built on the device."""
        return cls(A, b, None, m, n, _aux=True, **kw)

    @classmethod
    def synthetic_dense(cls, m, n, seed=0, pos_permille=1000, **kw) -> "LPState":
        return cls.synthetic(N.LPS_GEN_DENSE, m, n, seed, pos_permille, **kw)

    @classmethod
    def synthetic(cls, kind, m, n, seed=0, param=1000, **kw) -> "LPState":
        """One of the synthetic families of SURVEY.md §8d, generated in HBM (lps_generate_lp)."""
        st = cls.__new__(cls)
        st._lib = N.load()
        st._h = c_void_p()
        st._names0 = None
        opts = N.default_options()
        opts.epsilon = kw.get("epsilon", cls.DEF_EPSILON)
        opts.inf = kw.get("inf", cls.DEF_INF)
        opts.device = kw.get("device", -1)
        opts.time_kernels = int(kw.get("time_kernels", False))
        opts.update_variant = int(kw.get("update_variant", -1))
        opts.loop_mode = int(kw.get("loop_mode", 0))
        opts.block_pivots = int(kw.get("block_pivots", 0))
        opts.panel_ctas = int(kw.get("panel_ctas", 0))
        opts.pass_chunk_rows = int(kw.get("pass_chunk_rows", 0))
        rc = st._lib.lps_create(byref(st._h), byref(opts))
        if rc != N.LPS_OK:
            raise LpsError(rc, "lps_create: " + st._lib.lps_status_string(rc).decode())
        st._ck(st._lib.lps_generate_lp(st._h, int(kind), m, n, seed, int(param)), "lps_generate_lp")
        st._synthetic = (int(kind), m, n, seed, int(param))
        return st

    def regenerate(self) -> None:
        """Generate the same synthetic LP again on this handle (state, positions and log start over)."""
        kind, m, n, seed, param = self._synthetic
        self._ck(self._lib.lps_generate_lp(self._h, kind, m, n, seed, param), "lps_generate_lp")

    def _ck(self, rc, what):
        if rc == N.LPS_OK:
            return
        msg = "%s: %s (%s)" % (what, self._lib.lps_status_string(rc).decode(),
                               self._lib.lps_last_error(self._h).decode())
        if rc == N.LPS_ERR_INVALID:
            raise ValueError(msg)      # IllegalArgumentException from Validate.isTrue, LPState.java:288
        raise LpsError(rc, msg)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.lps_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the reference's method set --------------------------------------------------------
    def get_entering(self) -> int:
        e = c_int(-1)
        self._ck(self._lib.lps_get_entering(self._h, byref(e)), "getEntering")
        return e.value

    def get_leaving(self, entering: int) -> int:
        l = c_int(-1)
        self._ck(self._lib.lps_get_leaving(self._h, int(entering), byref(l)), "getLeaving")
        return l.value

    def pivot(self, entering: int, leaving: int) -> None:
        self._ck(self._lib.lps_pivot(self._h, int(entering), int(leaving)), "pivot")

    getEntering, getLeaving = get_entering, get_leaving
    pivotSequentially = pivotConcurrently = pivot   # LPState.java:133,184: same values either way

    def run(self, max_pivots: int = -1) -> N.LpsRunResult:
        res = N.LpsRunResult()
        self._ck(self._lib.lps_run(self._h, int(max_pivots), byref(res)), "run")
        return res

    # -- fields -------------------------------------------------------------------------
    def _dims(self) -> Tuple[int, int]:
        m, n = c_int(), c_int()
        self._ck(self._lib.lps_dims(self._h, byref(m), byref(n)), "dims")
        return m.value, n.value

    @property
    def m(self) -> int:
        return self._dims()[0]

    @property
    def n(self) -> int:
        return self._dims()[1]

    @property
    def v(self) -> float:
        v = c_double()
        self._ck(self._lib.lps_read_v(self._h, byref(v)), "read_v")
        return v.value

    @property
    def b(self) -> np.ndarray:
        out = np.empty(self.m, dtype=np.float64)
        if out.size:
            self._ck(self._lib.lps_read_b(self._h, _dp(out)), "read_b")
        return out

    @property
    def c(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.float64)
        if out.size:
            self._ck(self._lib.lps_read_c(self._h, _dp(out)), "read_c")
        return out

    @property
    def A(self) -> np.ndarray:
        m, n = self._dims()
        out = np.empty((m, n), dtype=np.float64)
        if out.size:
            self._ck(self._lib.lps_read_A(self._h, _dp(out), n), "read_A")
        return out

    def row(self, i: int) -> np.ndarray:
        out = np.empty(self.n, dtype=np.float64)
        self._ck(self._lib.lps_read_row(self._h, int(i), _dp(out)), "read_row")
        return out

    def col(self, j: int) -> np.ndarray:
        out = np.empty(self.m, dtype=np.float64)
        self._ck(self._lib.lps_read_col(self._h, int(j), _dp(out)), "read_col")
        return out

    @property
    def positions(self) -> np.ndarray:
        m, n = self._dims()
        out = np.empty(m + n, dtype=np.int32)
        if out.size:
            self._ck(self._lib.lps_read_positions(self._h, _ip(out)), "read_positions")
        return out

    def position_of(self, var: int) -> int:
        p = c_int(-1)
        self._ck(self._lib.lps_position_of(self._h, int(var), byref(p)), "position_of")
        return p.value

    @property
    def pivot_log(self) -> List[Tuple[int, int]]:
        cnt = c_int64(0)
        self._ck(self._lib.lps_read_pivot_log(self._h, None, 0, byref(cnt)), "pivot_log")
        buf = np.zeros((max(cnt.value, 1), 2), dtype=np.int32)
        if cnt.value:
            self._ck(self._lib.lps_read_pivot_log(self._h, _ip(buf), cnt.value, byref(cnt)), "pivot_log")
        return [(int(e), int(l)) for e, l in buf[:cnt.value]]

    def primal(self, nvars: int) -> np.ndarray:
        x = np.zeros(nvars, dtype=np.float64)
        if nvars:
            self._ck(self._lib.lps_read_primal(self._h, nvars, _dp(x)), "read_primal")
        return x

    # name maps, derived from the permutation (LPState.java:27-28, :311-320)
    @property
    def variables(self) -> Optional[Dict[int, str]]:
        if self._names0 is None:
            return None
        return {pos: self._names0[var] for pos, var in enumerate(self.positions)}

    @property
    def coefficients(self) -> Optional[Dict[str, int]]:
        if self._names0 is None:
            return None
        return {self._names0[var]: pos for pos, var in enumerate(self.positions)}

    def has_variables_names(self) -> bool:
        return self._names0 is not None

    # phase-1 support ---------------------------------------------------------------------
    def first_nonzero_in_row(self, row: int) -> int:
        j = c_int(-1)
        self._ck(self._lib.lps_first_nonzero_in_row(self._h, int(row), byref(j)), "first_nonzero_in_row")
        return j.value

    def drop_column(self, j: int) -> None:
        self._ck(self._lib.lps_drop_column(self._h, int(j)), "drop_column")

    def rebuild_objective(self, ops) -> None:
        arr = (N.LpsObjectiveOp * max(len(ops), 1))()
        for k, (kind, index, coef) in enumerate(ops):
            arr[k].kind, arr[k].index, arr[k].coef = int(kind), int(index), float(coef)
        self._ck(self._lib.lps_rebuild_objective(self._h, arr, len(ops)), "rebuild_objective")

    def measure_fp64_issue_rate(self, ms: float = 100.0) -> float:
        """FP64 thread-instructions per second of the pass's DMUL + DADD mix on this GPU (roofline aid)."""
        x = c_double()
        self._ck(self._lib.lps_measure_fp64_issue_rate(self._h, float(ms), byref(x)), "measure_fp64_issue_rate")
        return x.value

    def loop_description(self) -> str:
        """which kernels lps_run uses for this LP (loop shape, pass kernel, SM split)"""
        buf = ctypes.create_string_buffer(512)
        self._ck(self._lib.lps_loop_description(self._h, buf, 512), "loop_description")
        return buf.value.decode()

    def algorithmic_bytes_per_pivot(self) -> int:
        x = c_int64()
        self._ck(self._lib.lps_algorithmic_bytes_per_pivot(self._h, byref(x)), "bytes")
        return x.value
