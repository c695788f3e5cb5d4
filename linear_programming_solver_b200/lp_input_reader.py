"""`LPInputReader` (LPInputReader.java:19-224): the reference's text format, parsed by the native
host layer (csrc/lp_input_reader.cpp) — no JVM and no GPU needed.

    max
    2x1 + 3.05*x3
    1.05*x4 + 25*x1 == 0
    x1 + x2 + x3 + x24 >= 0
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_void_p

import numpy as np

from . import _native as N
from .exceptions import LPException
from .lp_standard_form import LPStandardForm


class LPInputReader:
    def _read(self, text: str, file_semantics: bool) -> LPStandardForm:
        if text is None:
            raise ValueError("null")          # IllegalArgumentException from @NotNull (LPInputReader.java:52,96)
        lib = N.load()
        m, n, mx = c_int(), c_int(), c_int()
        A, b, c = POINTER(c_double)(), POINTER(c_double)(), POINTER(c_double)()
        names = c_void_p()
        err = ctypes.create_string_buffer(256)
        rc = lib.lpsolver_read_lp(text.encode(), int(file_semantics), byref(m), byref(n), byref(mx), byref(A), byref(b),
                                  byref(c), byref(names), err, 256)
        if rc == 1:
            raise LPException(err.value.decode())
        if rc != 0:
            raise ValueError(err.value.decode())   # NumberFormatException and friends
        try:
            mm, nn = m.value, n.value
            An = np.ctypeslib.as_array(A, shape=(max(mm * nn, 1),))[: mm * nn].copy().reshape(mm, nn)
            bn = np.ctypeslib.as_array(b, shape=(max(mm, 1),))[:mm].copy()
            cn = np.ctypeslib.as_array(c, shape=(max(nn, 1),))[:nn].copy()
            name_list = ctypes.string_at(names).decode().split("\n") if nn else []
        finally:
            for p in (A, b, c):
                lib.lpsolver_free(ctypes.cast(p, c_void_p))
            lib.lpsolver_free(names)
        variables = {i: nm for i, nm in enumerate(name_list)}
        coefficients = {}
        for i, nm in enumerate(name_list):
            coefficients[nm] = i
        return LPStandardForm(An, bn, cn, mm, nn, bool(mx.value), variables, coefficients)

    def read_lp(self, lp: str) -> LPStandardForm:
        """readLP(String), LPInputReader.java:96-114."""
        return self._read(lp, False)

    def read_lp_file(self, path) -> LPStandardForm:
        """readLP(File), LPInputReader.java:52-93: stops at the first blank line after the constraints."""
        import os
        if path is None or not os.path.isfile(path) or not os.access(path, os.R_OK):
            raise ValueError("not a readable file")      # IllegalArgumentException, :54-61
        with open(path) as f:
            return self._read(f.read(), True)

    readLP = read_lp
