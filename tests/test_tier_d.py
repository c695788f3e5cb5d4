"""The C decimal-15 oracle (oracle/tier_d.c) against Python's `decimal` (exact spec semantics)
operation by operation, and against the Python restatement on whole solves.  CPU only."""
import decimal
import random
from decimal import Decimal

import numpy as np
import pytest

from oracle import tier_d, tier_f
from oracle.arith import Dec15
from oracle.simplex_ref import LPSolver, LPStandardForm


def _rand_dec(rng, max_digits=26):
    nd = rng.choice([1, 2, 5, 14, 15, 16, 17, 21, max_digits])
    coef = rng.randrange(0, 10 ** nd)
    if rng.random() < 0.15:
        coef = int("5" + "0" * (nd - 1)) if nd > 1 else 5         # tie patterns
    if rng.random() < 0.1:
        coef = int(str(rng.randrange(1, 10 ** 15)) + "5" + "0" * rng.randrange(0, 8))
    if rng.random() < 0.1:
        coef = int(str(rng.randrange(1, 10 ** 15)) + "49999999"[: rng.randrange(1, 9)])
    if rng.random() < 0.05:
        coef = 10 ** rng.randrange(0, 20)
    if rng.random() < 0.05:
        coef = 10 ** 15 - 1
    exp = rng.choice([0, -1, -9, -20, 3, rng.randrange(-60, 60)])
    sign = rng.random() < 0.5
    return Decimal((int(sign), tuple(int(ch) for ch in str(coef)), exp)) if coef else Decimal(0)


def _same(a: Decimal, b: Decimal) -> bool:
    return (a == b) and not (a.is_nan() or b.is_nan())


def test_scalar_ops_match_python_decimal():
    rng = random.Random(12345)
    bad = []
    for k in range(60000):
        a, b = _rand_dec(rng), _rand_dec(rng)
        if rng.random() < 0.2:          # near-cancellation / close exponents
            b = a + _rand_dec(rng, 5) * Decimal(10) ** (a.adjusted() - rng.randrange(10, 20)) if a else b
            if len(b.as_tuple().digits) > 30:
                b = Dec15.add(b, Decimal(0))
        for name, fn in (("mul", Dec15.mul), ("add", Dec15.add), ("sub", Dec15.sub), ("div", Dec15.div)):
            if name == "div" and not b:
                continue
            if name == "mul" and len(a.as_tuple().digits) + len(b.as_tuple().digits) > 70:
                continue
            want = fn(a, b)
            got = tier_d.op(name, a, b)
            if not _same(want, got):
                bad.append((name, a, b, want, got))
        if tier_d.cmp(a, b) != Dec15.cmp(a, b):
            bad.append(("cmp", a, b, Dec15.cmp(a, b), tier_d.cmp(a, b)))
    assert not bad, bad[:5]


def test_far_apart_addends_and_ties():
    cases = [
        (Decimal("123456789012345"), Decimal("1E-60")),           # tiny addend cannot change 15 digits
        (Decimal("1234567890123455"), Decimal("-1E-60")),         # ...but breaks an exact tie downwards
        (Decimal("1234567890123455"), Decimal("1E-60")),
        (Decimal("1234567890123445"), Decimal("9.99E-40")),
        (Decimal("999999999999999"), Decimal("0.5")),             # carry into a 16th digit
        (Decimal("9999999999999995"), Decimal("0")),
        (Decimal("1E+50"), Decimal("-1E-9")),
        (Decimal("0.1"), Decimal("-0.1")),
    ]
    for a, b in cases:
        for name, fn in (("add", Dec15.add), ("sub", Dec15.sub)):
            assert _same(fn(a, b), tier_d.op(name, a, b)), (name, a, b)
            assert _same(fn(b, a), tier_d.op(name, b, a)), (name, b, a)


def test_from_double_is_exact():
    rng = random.Random(7)
    for _ in range(3000):
        x = rng.randrange(-10 ** 6, 10 ** 6) / 2 ** rng.randrange(0, 21)
        assert tier_d.from_double(x) == Decimal(x)
    assert tier_d.from_double(0.0) == 0
    with pytest.raises(ValueError):
        tier_d.from_double(0.1)      # 55 significant digits


@pytest.mark.parametrize("m,n,seed", [(5, 7, 0), (12, 9, 1), (20, 20, 2), (30, 45, 3), (40, 25, 4)])
def test_whole_solve_equals_python_tier_d(m, n, seed):
    A, b, c = tier_f.gen_dense_feasible(m, n, seed)
    form = LPStandardForm(A.tolist(), b.tolist(), c.tolist(), m, n, True, arith=Dec15)
    solver = LPSolver(Dec15)
    solver.solve(form)
    fs = solver.trace.final_state
    st = tier_d.TierDState(A, b, c, nthreads=2)
    status, k = st.run()
    assert status == tier_d.OPTIMAL and st.log == solver.trace.phase2_log
    assert st.v == fs.v
    for i in range(m):
        for j in range(n):
            assert st.cell(i, j) == fs.A[i][j]
        assert st.cell(i, n) == fs.b[i]
    for j in range(n):
        assert st.cell(m, j) == fs.c[j]


def test_selection_rules_and_spock_pivot():
    from tests.golden import spock_vectors as G
    v = G.PIVOT_4x5
    for case in v["cases"]:
        st = tier_d.TierDState(np.array(v["A"], dtype=float), np.array(v["b"], dtype=float), np.array(v["c"], dtype=float))
        st.pivot(case["e"], case["l"])
        A, b, c, vv, pos = st.read()
        assert A.tolist() == [[float(x) for x in r] for r in case["resA"]]
        assert b.tolist() == [float(x) for x in case["resB"]] and c.tolist() == [float(x) for x in case["resC"]]
        assert vv == float(case["resV"])
    g = G.GET_LEAVING                                   # 0.1 is not dyadic: use 0.125 in its place
    A = np.array([[0.125 if x == "0.1" else float(x) for x in r] for r in g["A"]])
    st = tier_d.TierDState(A, np.array(g["b"], dtype=float), np.zeros(4))
    assert [st.get_leaving(e) for e in range(4)] == [0, 3, 1, 3]
    st = tier_d.TierDState(np.zeros((0, 5)), np.zeros(0), np.array([0, 0, 0, 0, 1.0]))
    assert st.get_entering() == 4
