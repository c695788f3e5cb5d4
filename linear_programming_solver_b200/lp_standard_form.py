"""`LPStandardForm` (LPStandardForm.java:10-65): holder of A, b, c, m, n, maximize and the
optional name maps.  Numbers are binary64; anything float() accepts (int, str, Decimal) is
converted on entry (the reference keeps BigDecimal; see DESIGN.md on the number system)."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np


class LPStandardForm:
    def __init__(self, A, b, c, m: int, n: int, maximize: bool,
                 variables: Optional[Dict[int, str]] = None,
                 coefficients: Optional[Dict[str, int]] = None):
        self.A = np.ascontiguousarray(np.asarray(A, dtype=np.float64).reshape(m, n)) if m * n else np.zeros((m, n))
        self.b = np.array([float(x) for x in b], dtype=np.float64) if not isinstance(b, np.ndarray) else b.astype(np.float64)
        self.c = np.array([float(x) for x in c], dtype=np.float64) if not isinstance(c, np.ndarray) else c.astype(np.float64)
        self.m = m
        self.n = n
        self.maximize = maximize
        self.variables = variables
        self.coefficients = coefficients

    def has_variable_names(self) -> bool:  # LPStandardForm.java:154-156
        return self.variables is not None and self.coefficients is not None

    hasVariableNames = has_variable_names

    def get_dual(self) -> "LPStandardForm":
        """LPStandardForm.getDual (LPStandardForm.java:129-152)."""
        B = np.ascontiguousarray(self.A.T)
        variables = coefficients = None
        if self.has_variable_names():
            variables = {i: "x%d" % (i + 1) for i in range(self.n)}
            coefficients = {"x%d" % (i + 1): i for i in range(self.n)}
        return LPStandardForm(B, self.c.copy(), self.b.copy(), self.n, self.m, not self.maximize,
                              variables, coefficients)

    getDual = get_dual
