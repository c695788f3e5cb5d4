"""ctypes front end of the C decimal-15 oracle (`tier_d.c`) — TEST INFRASTRUCTURE (oracle/__init__.py).

`TierDState` is LPState (LPState.java:17-320) in the reference's own arithmetic —
BigDecimal + MathContext(15, HALF_UP) — at C speed, so parity against the reference's number
system can be checked at BASELINE sizes (hundreds to a thousand rows) instead of only on tiny LPs.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from decimal import Decimal
from typing import List, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
OPTIMAL, UNBOUNDED, PIVOT_CAP = 0, 1, 2


class DecIO(ctypes.Structure):
    _fields_ = [("lo", ctypes.c_uint64), ("hi", ctypes.c_uint64), ("exp", ctypes.c_int32), ("neg", ctypes.c_int32)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libtier_d.so")
    src = os.path.join(_HERE, "tier_d.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libtier_d.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        pio = ctypes.POINTER(DecIO)
        L.td_op.argtypes = [ctypes.c_int, pio, pio, pio]
        L.td_op.restype = None
        L.td_cmp.argtypes = [pio, pio]
        L.td_cmp.restype = ctypes.c_int
        L.td_from_double.argtypes = [ctypes.c_double, pio]
        L.td_from_double.restype = ctypes.c_int
        L.td_create.argtypes = [ctypes.c_int, ctypes.c_int, dp, ctypes.c_long, dp, dp]
        L.td_create.restype = ctypes.c_void_p
        L.td_destroy.argtypes = [ctypes.c_void_p]
        L.td_get_entering.argtypes = [ctypes.c_void_p]
        L.td_get_entering.restype = ctypes.c_int
        L.td_get_leaving.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.td_get_leaving.restype = ctypes.c_int
        L.td_pivot.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.td_pivot.restype = None
        L.td_run.argtypes = [ctypes.c_void_p, ctypes.c_long, ip, ctypes.c_long, ctypes.POINTER(ctypes.c_long),
                             ctypes.c_int]
        L.td_run.restype = ctypes.c_int
        L.td_read.argtypes = [ctypes.c_void_p, dp, dp, dp, dp, ip]
        L.td_read.restype = None
        L.td_v_string.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
        L.td_cell_string.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
        _LIB = L
    return _LIB


def to_io(d: Decimal) -> DecIO:
    sign, digits, exp = d.as_tuple()
    coef = int("".join(map(str, digits))) if digits else 0
    assert coef < 10 ** 38
    return DecIO(coef & (2 ** 64 - 1), coef >> 64, exp, 1 if (sign and coef) else 0)


def from_io(x: DecIO) -> Decimal:
    coef = (x.hi << 64) | x.lo
    return Decimal((1 if x.neg else 0, tuple(int(ch) for ch in str(coef)), x.exp))


OPS = {"mul": 0, "add": 1, "sub": 2, "div": 3}


def op(name: str, a: Decimal, b: Decimal) -> Decimal:
    r = DecIO()
    lib().td_op(OPS[name], ctypes.byref(to_io(a)), ctypes.byref(to_io(b)), ctypes.byref(r))
    return from_io(r)


def cmp(a: Decimal, b: Decimal) -> int:
    return lib().td_cmp(ctypes.byref(to_io(a)), ctypes.byref(to_io(b)))


def from_double(x: float) -> Decimal:
    r = DecIO()
    rc = lib().td_from_double(x, ctypes.byref(r))
    if rc:
        raise ValueError("binary64 value needs more than 38 decimal digits")
    return from_io(r)


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class TierDState:
    def __init__(self, A, b, c, nthreads: int = 1):
        A = np.ascontiguousarray(A, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        self.m, self.n = len(b), len(c)
        self.nthreads = nthreads
        self._h = lib().td_create(self.m, self.n, _dp(A), self.n, _dp(b), _dp(c))
        if not self._h:
            raise ValueError("an input needs more than 38 decimal digits")
        self.log: List[Tuple[int, int]] = []

    def __del__(self):
        if getattr(self, "_h", None):
            lib().td_destroy(self._h)
            self._h = None

    def get_entering(self) -> int:
        return lib().td_get_entering(self._h)

    def get_leaving(self, e: int) -> int:
        return lib().td_get_leaving(self._h, e)

    def pivot(self, e: int, l: int) -> None:
        lib().td_pivot(self._h, e, l, self.nthreads)
        self.log.append((e, l))

    def run(self, max_pivots: int = -1, log_cap: int = 1 << 22):
        cap = log_cap if max_pivots < 0 else min(log_cap, max_pivots)
        buf = np.zeros((max(cap, 1), 2), dtype=np.int32)
        k = ctypes.c_long(0)
        status = lib().td_run(self._h, max_pivots, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), cap,
                              ctypes.byref(k), self.nthreads)
        self.log.extend((int(e), int(l)) for e, l in buf[:min(k.value, cap)])
        return status, k.value

    def read(self):
        """(A, b, c, v, pos2var) with every value converted to the nearest binary64."""
        A = np.empty((self.m, self.n))
        b = np.empty(self.m)
        c = np.empty(self.n)
        v = ctypes.c_double()
        pos = np.empty(self.m + self.n, dtype=np.int32)
        lib().td_read(self._h, _dp(A), _dp(b), _dp(c), ctypes.byref(v), pos.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
        return A, b, c, v.value, pos

    @property
    def v(self) -> Decimal:
        buf = ctypes.create_string_buffer(96)
        lib().td_v_string(self._h, buf, 96)
        return Decimal(buf.value.decode())

    def cell(self, i: int, j: int) -> Decimal:
        """A[i][j] (j < n), b[i] (j == n) or c[j] (i == m) as an exact Decimal."""
        buf = ctypes.create_string_buffer(96)
        lib().td_cell_string(self._h, i, j, buf, 96)
        return Decimal(buf.value.decode())
