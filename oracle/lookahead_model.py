"""CPU model of the LOOK-AHEAD blocked loop (csrc/lps_step.cuh) — test infrastructure, not product.

One launch of the product does two things at once: the pass applies the pending set q to the tableau
out of place (T_next <- T_cur with block k applied) while the panel decides the pivots of block k+1.
The panel cannot read "T after block k" — nobody has written it yet — so it reads T_cur and replays
block k's pivots and then its own on every cell it touches.  It also stops re-deriving the b column
and the objective row for every pivot: both are carried as running vectors, updated by the one
multiply-subtract per entry the reference applies (LPState.java:164 for b, :177 for c; the leaving
row's b entry and the entering column's c entry as :146 / :172 set them).

This module restates that schedule with numpy (binary64, every multiply / subtract / divide rounded on
its own) so that the two claims the CUDA code rests on are checked on the CPU against Tier F, bit for
bit (tests/test_lookahead_model.py):
  1. replaying [previous block, own block] on T_cur yields the cells of the current state;
  2. the running vectors equal the b column / objective row a replay would produce — in particular the
     ratio test (LPState.java:287-305) and the entering rule (:274-285) see the same bits.
The launch structure is kept: `launches` counts them (ceil(pivots / block) + 1 for a capped run).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

OPTIMAL, UNBOUNDED, PIVOT_CAP = 0, 1, 2


class _Set:
    """one set of pending pivots (CtlS::blk_e2 / blk_l2 / blk_p2 + its pending columns and rows)"""

    def __init__(self):
        self.e: List[int] = []
        self.l: List[int] = []
        self.p: List[float] = []
        self.a: List[np.ndarray] = []     # OLD entering column, rows 0..m (objective row included)
        self.r: List[np.ndarray] = []     # NEW (scaled) pivot row, columns 0..n

    def __len__(self):
        return len(self.e)

    def items(self):
        return zip(self.e, self.l, self.p, self.a, self.r)


class LookAheadModel:
    def __init__(self, A, b, c, v=0.0, block=16, eps=1e-9, inf=1e50):
        A = np.asarray(A, dtype=np.float64)
        self.m, self.n = A.shape
        m, n = self.m, self.n
        T = np.zeros((m + 1, n + 1))
        T[:m, :n] = A
        T[:m, n] = b
        T[m, :n] = c
        T[m, n] = -v
        self.Tbuf = [T, np.zeros_like(T)]      # lps_handle::T / T2
        self.cur = 0                           # CtlS::cur_at
        self.block, self.eps, self.inf = block, eps, inf
        self.log: List[Tuple[int, int]] = []
        self.launches = 0
        self.passes = 0

    # -- the pass role (sweep_role / flush_role): set q on every cell, out of place ------------------
    def _pass(self, pend: _Set) -> None:
        if not len(pend):
            return
        src = self.Tbuf[self.cur]
        dst = self.Tbuf[self.cur ^ 1]
        np.copyto(dst, src)
        for e, l, p, a, r in pend.items():
            col_e = -(a / p)
            dst -= np.multiply.outer(a, r)
            dst[:, e] = col_e
            if l >= 0:
                dst[l, :] = r
        self.cur ^= 1
        self.passes += 1

    # -- the panel role's lazy evaluation: T_cur, then the previous block, then my own ---------------
    @staticmethod
    def _replay_column(x, j, sets):
        for s in sets:
            for e, l, p, a, r in s.items():
                if j == e:
                    x = -(a / p)
                else:
                    x = x - a * r[j]
                if l >= 0:                  # (-1: the leaving row lives on another rank)
                    x[l] = r[j]
        return x

    @staticmethod
    def _replay_row(x, i, sets):
        for s in sets:
            for e, l, p, a, r in s.items():
                if l >= 0 and i == l:
                    x = r.copy()
                else:
                    xe = -(a[i] / p)
                    x = x - a[i] * r
                    x[e] = xe
        return x

    # -- the two per-pivot decisions; a sharded rank overrides them with the exchanges --------------
    def _leaving(self, a, bvec):
        """ratio test over my rows on the RUNNING b (LPState.java:287-305) -> (row, pivot element)"""
        best, l = self.inf, -1
        ok = ~(a[:self.m] < self.eps)
        with np.errstate(all="ignore"):
            for i in np.nonzero(ok)[0]:
                s = bvec[i] / a[i]
                if s < best:
                    best, l = s, int(i)
        return best, l

    def _decide(self, e, a, bvec):
        """-> (global leaving row or -1, pivot element, my local index of that row or -1)"""
        _, l = self._leaving(a, bvec) if e >= 0 else (self.inf, -1)
        return l, (a[l] if l >= 0 else 0.0), l

    def _pivot_row(self, Tread, l, lloc, p, e, sets):
        """the scaled leaving row (LPState.java:137-146): replayed and scaled where the row lives
        (l: global row, lloc: my local index of it)"""
        r = self._replay_row(Tread[lloc, :].copy(), lloc, sets) / p
        r[e] = 1.0 / p
        return r

    def run(self, max_pivots: int = -1):
        """lps_run with loop_mode 7.  Returns (verdict, pivots)."""
        m, n = self.m, self.n
        # kb_init_vec: the tableau is fully applied between runs
        T0 = self.Tbuf[self.cur]
        bvec = T0[:, n].copy()                 # rows 0..m (slot m: -v)
        cvec = T0[m, :].copy()                 # columns 0..n (slot n: -v)
        pos = np.nonzero(cvec[:n] > self.eps)[0]
        e = int(pos[0]) if pos.size else -1    # ks_first_positive
        prev, done, verdict = _Set(), 0, None
        while True:
            # ---- one launch: pass(prev) beside panel(next block) --------------------------------
            self.launches += 1
            Tread = self.Tbuf[self.cur]        # both roles read it; the pass writes the other buffer
            own = _Set()
            while verdict is None and len(own) < self.block:
                a = self._replay_column(Tread[:, e].copy(), e, (prev, own)) if e >= 0 else None      # phase A
                l, p, lloc = self._decide(e, a, bvec)
                if e < 0:
                    verdict = OPTIMAL
                    break
                if l < 0:
                    verdict = UNBOUNDED
                    break
                if 0 <= max_pivots <= done:
                    verdict = PIVOT_CAP
                    break
                r = self._pivot_row(Tread, l, lloc, p, e, (prev, own))                                    # phase B
                # running vectors: one multiply-subtract per entry (LPState.java:164 / :177)
                rn, ce = r[n], a[m]
                nb = bvec - a * rn
                if lloc >= 0:
                    nb[lloc] = rn
                nc = cvec - ce * r
                nc[e] = -(ce / p)
                # the objective slot is the same cell in both vectors: (m, n)
                assert nb[m].tobytes() == nc[n].tobytes()
                bvec, cvec = nb, nc
                own.e.append(e); own.l.append(lloc); own.p.append(p); own.a.append(a); own.r.append(r)
                self.log.append((e, l))
                done += 1
                pos = np.nonzero(cvec[:n] > self.eps)[0]
                e = int(pos[0]) if pos.size else -1
            self._pass(prev)                   # (concurrently, on the other SMs)
            prev = own
            if verdict is not None and not len(prev):
                break
        # what the claims say: the running vectors ARE the tableau's last column / row
        T = self.Tbuf[self.cur]
        assert np.array_equal(bvec, T[:, n]) and np.array_equal(cvec, T[m, :])
        return verdict, done

    # -- views -------------------------------------------------------------------------------------
    @property
    def T(self):
        return self.Tbuf[self.cur]

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


class ShardedLookAheadModel(LookAheadModel):
    """One rank of the row-sharded look-ahead loop (kb_step<true, ...>): rows [lo, hi) of (A | b) plus a
    replica of the objective row; the running b is local, the running c replicated.  Two callbacks stand
    for what crosses NVLink per pivot in the product: `all_gather(obj) -> list` (the ratio-test candidates,
    CTA 0's packets) and `broadcast(obj, src) -> obj` (the scaled pivot row, the owner's packets)."""

    def __init__(self, A_local, b_local, c, lo, hi, m_total, owner_of, all_gather, broadcast, rank,
                 v=0.0, block=16, eps=1e-9, inf=1e50):
        super().__init__(A_local, b_local, c, v=v, block=block, eps=eps, inf=inf)
        self.lo, self.hi, self.m_total = lo, hi, m_total
        self.owner_of, self.all_gather, self.broadcast, self.rank = owner_of, all_gather, broadcast, rank

    def _decide(self, e, a, bvec):
        best, l_loc = self._leaving(a, bvec) if e >= 0 else (self.inf, -1)
        cands = self.all_gather((float(best), self.lo + l_loc if l_loc >= 0 else -1,
                                 float(a[l_loc]) if l_loc >= 0 else 0.0))
        win = (self.inf, -1, 0.0)
        for cand in cands:                    # lexicographic (ratio, global row): lowest row wins ties
            if cand[1] >= 0 and (win[1] < 0 or cand[0] < win[0] or (cand[0] == win[0] and cand[1] < win[1])):
                win = cand
        l, p = win[1], win[2]
        mine = l >= 0 and self.owner_of(l) == self.rank
        return l, p, (l - self.lo if mine else -1)

    def _pivot_row(self, Tread, l, lloc, p, e, sets):
        r = super()._pivot_row(Tread, l, lloc, p, e, sets) if lloc >= 0 else None      # compute + broadcast on the owner
        return np.asarray(self.broadcast(r, self.owner_of(l)))
