"""Restatement of `LPInputReader` (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows LPInputReader.java:19-224: the three regexes (:25-31), `readLP(String)` (:96-114),
`readLP(File)` semantics of stopping at the first blank line after at least one constraint
(:76-86), objective tokenisation (:131-155), constraint lowering — `>=` rows negated, `=` /
`==` rows emitted as a +/- pair (:189-212) — and zero padding of short rows (:215-223).
"""
from __future__ import annotations

import re
from decimal import Decimal
from typing import List

from .arith import Dec15
from .simplex_ref import LPException, LPStandardForm

_A = re.ASCII
_OBJECTIVE = re.compile(r"^((\s*[+-]?\s*\d*\.?\d*)\*?([a-zA-Z]+\d*))+\s*$", _A)
_CONSTRAINT = re.compile(
    r"^((\s*[+-]?\s*\d*\.?\d*)\*?([a-zA-Z]+\d*))+\s*(=|==|<=|>=)\s*(-?\s*\d+(\.\d+)?)\s*$", _A)
_TOKEN = re.compile(r"(([+-]?\s*\d*\.?\d*)\*?([a-zA-Z]+\d*))", _A)


def _neg(x: Decimal) -> Decimal:
    return -x if x else abs(x)  # BigDecimal has no signed zero


class LPInputReader:
    def __init__(self, arith=Dec15):
        self.arith = arith

    def _reload(self):
        self.A: List[List[Decimal]] = []
        self.b: List[Decimal] = []
        self.c: List[Decimal] = []
        self.variables = {}
        self.coefficients = {}
        self.num_vars = 0
        self.num_ineq = 0

    def read_lp(self, text: str) -> LPStandardForm:
        """readLP(String), LPInputReader.java:96-114."""
        self._reload()
        lines = text.split("\n")
        while lines and lines[-1] == "":      # java.lang.String.split drops trailing empty strings
            lines.pop()
        if len(lines) < 3:
            raise LPException("Incomplete lp")
        maximize = self._max_min(lines[0])
        self.c = self._objective(lines[1])
        for line in lines[2:]:
            self._constraint(line)
        return self._finish(maximize)

    def read_lp_file_text(self, text: str) -> LPStandardForm:
        """readLP(File), LPInputReader.java:52-93, applied to the file's text: first line
        max/min, second the objective, then constraints until the first blank line."""
        self._reload()
        lines = text.splitlines()
        if not lines:
            raise LPException("Input file is empty")
        maximize = self._max_min(lines[0])
        self.c = self._objective(lines[1])
        count = 0
        for line in lines[2:]:
            if line.strip():
                self._constraint(line)
                count += 1
            elif count > 0:
                break
            else:
                raise LPException("No constraints in the input file")
        return self._finish(maximize)

    def _finish(self, maximize):
        for row in self.A:                                   # normalizeConstraintMatrix :215-223
            row.extend([Decimal(0)] * (self.num_vars - len(row)))
        cv = self.arith.conv
        A = [[cv(x) for x in row] for row in self.A]
        return LPStandardForm(A, [cv(x) for x in self.b], [cv(x) for x in self.c], self.num_ineq,
                              self.num_vars, maximize, self.variables, self.coefficients,
                              arith=self.arith)

    @staticmethod
    def _max_min(s: str) -> bool:                             # :117-128
        s = s.strip().lower()
        if s == "min":
            return False
        if s == "max":
            return True
        raise LPException("Incorrect max/min parameter")

    @staticmethod
    def _coef(t: str) -> Decimal:
        t = re.sub(r"\s+", "", t, flags=_A)
        if t in ("", "+"):
            t = "1"
        elif t == "-":
            t = "-1"
        return Decimal(t)

    def _objective(self, objective: str) -> List[Decimal]:    # :131-155
        if not _OBJECTIVE.fullmatch(objective):
            raise LPException("Can't recognize objective")
        out = []
        for i, mt in enumerate(_TOKEN.finditer(objective)):
            name = mt.group(3)
            self.variables[i] = name
            self.coefficients[name] = i
            out.append(self._coef(mt.group(2)))
        self.num_vars = len(self.variables)
        return out

    def _constraint(self, constraint: str) -> None:           # :158-213
        cm = _CONSTRAINT.search(constraint)
        if not cm:
            raise LPException("Can't recognize constraint")
        coefs = [Decimal(0)] * self.num_vars
        for mt in _TOKEN.finditer(constraint):
            name = mt.group(3)
            if name not in self.coefficients:
                self.variables[self.num_vars] = name
                self.coefficients[name] = self.num_vars
                self.num_vars += 1
                coefs.append(None)
                self.c.append(Decimal(0))
            coefs[self.coefficients[name]] = self._coef(mt.group(2))
        sign = cm.group(4).strip()
        rhs = Decimal(re.sub(r"\s", "", cm.group(5), flags=_A))
        if sign == ">=":
            self.A.append([_neg(x) for x in coefs])
            self.b.append(_neg(rhs))
            self.num_ineq += 1
        elif sign in ("==", "="):
            self.A.append(coefs)
            self.A.append([_neg(x) for x in coefs])
            self.b.append(rhs)
            self.b.append(_neg(rhs))
            self.num_ineq += 2
        else:
            self.A.append(coefs)
            self.b.append(rhs)
            self.num_ineq += 1


def split_lp_file(text: str) -> List[str]:
    """Split a multi-LP file such as io_files/input.txt into its blank-line-separated LPs."""
    blocks, cur = [], []
    for line in text.splitlines():
        if line.strip():
            cur.append(line)
        elif cur:
            blocks.append("\n".join(cur))
            cur = []
    if cur:
        blocks.append("\n".join(cur))
    return blocks
