"""CPU model of the BLOCKED pivot loop (csrc/lps_blocked.cuh) — test infrastructure, not product.

The product defers up to `block` pivots and applies them in one pass over the tableau.  Whatever it
needs before the pass — the entering column and the b column for LPState.getLeaving
(LPState.java:287-305), the leaving row to scale (LPState.java:137-146), the objective row for
LPState.getEntering (LPState.java:274-285) — it evaluates lazily by REPLAYING the pending pivots on
the block-start tableau:

    x <- T[i][j];  for u in pending:   i == l_u  ->  x = r_u[j]
                                       j == e_u  ->  x = -(a_u[i] / p_u)
                                       else      ->  x = x - a_u[i] * r_u[j]

This module restates that algorithm with numpy (binary64, every multiply / subtract / divide
rounded on its own, as in oracle/tier_f.c) so that the claim "the blocked loop performs, on every
cell, exactly the operations of the pivot-per-pass loop" is checked on the CPU against Tier F —
pivot sequence and every tableau cell, bit for bit (tests/test_blocked_model.py).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

OPTIMAL, UNBOUNDED, PIVOT_CAP = 0, 1, 2


class BlockedModel:
    def __init__(self, A, b, c, v=0.0, block=16, eps=1e-9, inf=1e50):
        A = np.asarray(A, dtype=np.float64)
        self.m, self.n = A.shape
        m, n = self.m, self.n
        # augmented tableau [A | b ; c | -v], the layout of DESIGN.md §3
        self.T = np.zeros((m + 1, n + 1))
        self.T[:m, :n] = A
        self.T[:m, n] = b
        self.T[m, :n] = c
        self.T[m, n] = -v
        self.block, self.eps, self.inf = block, eps, inf
        self.pend_e: List[int] = []
        self.pend_l: List[int] = []
        self.pend_p: List[float] = []
        self.pend_a: List[np.ndarray] = []    # a_u: OLD entering column, all m+1 rows
        self.pend_r: List[np.ndarray] = []    # r_u: NEW (scaled) pivot row, all n+1 columns
        self.log: List[Tuple[int, int]] = []
        self.passes = 0

    # -- lazy evaluation -------------------------------------------------------------------------
    def _column(self, j: int) -> np.ndarray:
        """column j of the CURRENT state (kb_col / phase A of kb_panel)"""
        x = self.T[:, j].copy()
        for e, l, p, a, r in zip(self.pend_e, self.pend_l, self.pend_p, self.pend_a, self.pend_r):
            if j == e:
                x = -(a / p)
            else:
                x = x - a * r[j]
            x[l] = r[j]
        return x

    def _row(self, i: int) -> np.ndarray:
        """row i of the CURRENT state (kb_row / phase B of kb_panel)"""
        x = self.T[i, :].copy()
        for e, l, p, a, r in zip(self.pend_e, self.pend_l, self.pend_p, self.pend_a, self.pend_r):
            if i == l:
                x = r.copy()
            else:
                xe = -(a[i] / p)
                x = x - a[i] * r
                x[e] = xe
        return x

    def flush(self) -> None:
        """kb_flush: every pending pivot on every cell, one pass"""
        if not self.pend_e:
            return
        T = self.T
        for e, l, p, a, r in zip(self.pend_e, self.pend_l, self.pend_p, self.pend_a, self.pend_r):
            col_e = -(a / p)
            T -= np.multiply.outer(a, r)      # product rounded, then difference rounded: no FMA in numpy
            T[:, e] = col_e
            T[l, :] = r
        self.pend_e, self.pend_l, self.pend_p, self.pend_a, self.pend_r = [], [], [], [], []
        self.passes += 1

    # -- the loop (LPSolver.java:101-112) ----------------------------------------------------------
    def run(self, max_pivots: int = -1):
        m, n = self.m, self.n
        done = 0
        while True:
            crow = self._row(m)
            pos = np.nonzero(crow[:n] > self.eps)[0]
            if pos.size == 0:
                self.flush()
                return OPTIMAL, done
            e = int(pos[0])
            a = self._column(e)
            bcol = self._column(n)
            best, l = self.inf, -1
            ok = ~(a[:m] < self.eps)
            ratios = np.full(m, self.inf)
            with np.errstate(all="ignore"):
                ratios[ok] = bcol[:m][ok] / a[:m][ok]
            for i in np.nonzero(ok)[0]:           # strict '<': the first row wins ties
                if ratios[i] < best:
                    best, l = ratios[i], int(i)
            if l < 0:
                self.flush()
                return UNBOUNDED, done
            if 0 <= max_pivots <= done:
                self.flush()
                return PIVOT_CAP, done
            p = a[l]
            r = self._row(l) / p
            r[e] = 1.0 / p
            self.pend_e.append(e)
            self.pend_l.append(l)
            self.pend_p.append(p)
            self.pend_a.append(a)
            self.pend_r.append(r)
            self.log.append((e, l))
            done += 1
            if len(self.pend_e) == self.block:
                self.flush()

    # -- views -------------------------------------------------------------------------------------
    @property
    def A(self):
        return self.T[:self.m, :self.n]

    @property
    def b(self):
        return self.T[:self.m, self.n]

    @property
    def c(self):
        return self.T[self.m, :self.n]

    @property
    def v(self):
        return 0.0 - self.T[self.m, self.n]


class ShardedBlockedModel(BlockedModel):
    """One rank of the row-sharded blocked loop (csrc/lps_blocked.cuh with world > 1): the rank holds
    rows [lo, hi) of (A | b) plus a replica of the objective row, keeps ITS entries of the pending
    columns and a full copy of the pending rows, and talks to the other ranks through two callbacks —
    `all_gather(obj) -> list` for the ratio-test candidates and `broadcast(obj, src) -> obj` for the
    scaled pivot row — which is exactly what crosses NVLink per pivot in the product."""

    def __init__(self, A_local, b_local, c, lo, hi, m_total, owner_of, all_gather, broadcast, rank,
                 v=0.0, block=16, eps=1e-9, inf=1e50):
        super().__init__(A_local, b_local, c, v=v, block=block, eps=eps, inf=inf)
        self.lo, self.hi, self.m_total = lo, hi, m_total
        self.owner_of, self.all_gather, self.broadcast, self.rank = owner_of, all_gather, broadcast, rank

    def run(self, max_pivots: int = -1):
        m, n = self.m, self.n                     # m = local rows; local row m is the objective replica
        done = 0
        while True:
            crow = self._row(m)
            pos = np.nonzero(crow[:n] > self.eps)[0]
            e = int(pos[0]) if pos.size else -1
            best, l_loc = self.inf, -1
            a = None
            if e >= 0:
                a = self._column(e)
                bcol = self._column(n)
                for i in range(m):
                    if not (a[i] < self.eps):
                        s = bcol[i] / a[i]
                        if s < best:
                            best, l_loc = s, i
            cands = self.all_gather((float(best), self.lo + l_loc if l_loc >= 0 else -1,
                                     float(a[l_loc]) if l_loc >= 0 else 0.0))
            if e < 0:
                self.flush()
                return OPTIMAL, done
            win = (self.inf, -1, 0.0)
            for cand in cands:                    # lexicographic (ratio, global row): lowest row wins ties
                if cand[1] >= 0 and (win[1] < 0 or cand[0] < win[0] or (cand[0] == win[0] and cand[1] < win[1])):
                    win = cand
            l, p = win[1], win[2]
            if l < 0:
                self.flush()
                return UNBOUNDED, done
            if 0 <= max_pivots <= done:
                self.flush()
                return PIVOT_CAP, done
            owner = self.owner_of(l)
            r = None
            if owner == self.rank:                # compute + broadcast: the owner replays and scales row l
                r = self._row(l - self.lo) / p
                r[e] = 1.0 / p
            r = self.broadcast(r, owner)
            self.pend_e.append(e)
            self.pend_l.append(l - self.lo if owner == self.rank else -1)
            self.pend_p.append(p)
            self.pend_a.append(a)
            self.pend_r.append(np.asarray(r))
            self.log.append((e, l))
            done += 1
            if len(self.pend_e) == self.block:
                self.flush()

    # a leaving row owned by another rank is "-1": it overwrites nothing here
    def _column(self, j):
        x = self.T[:, j].copy()
        for e, l, p, a, r in zip(self.pend_e, self.pend_l, self.pend_p, self.pend_a, self.pend_r):
            x = -(a / p) if j == e else x - a * r[j]
            if l >= 0:
                x[l] = r[j]
        return x

    def _row(self, i):
        x = self.T[i, :].copy()
        for e, l, p, a, r in zip(self.pend_e, self.pend_l, self.pend_p, self.pend_a, self.pend_r):
            if l >= 0 and i == l:
                x = r.copy()
            else:
                xe = -(a[i] / p)
                x = x - a[i] * r
                x[e] = xe
        return x

    def flush(self):
        if not self.pend_e:
            return
        T = self.T
        for e, l, p, a, r in zip(self.pend_e, self.pend_l, self.pend_p, self.pend_a, self.pend_r):
            col_e = -(a / p)
            T -= np.multiply.outer(a, r)
            T[:, e] = col_e
            if l >= 0:
                T[l, :] = r
        self.pend_e, self.pend_l, self.pend_p, self.pend_a, self.pend_r = [], [], [], [], []
        self.passes += 1
