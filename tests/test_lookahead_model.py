"""The look-ahead loop's schedule (csrc/lps_step.cuh: the panel of block k+1 replays [block k, own block]
on the tableau of BEFORE block k while the pass applies block k out of place; b column and objective row
carried as running vectors), restated on the CPU in oracle/lookahead_model.py, against the binary64
oracle: same pivot sequence, same verdict, every tableau cell bit-identical, for every block size.  The
GPU side of the same claim is tests/test_gpu_blocked.py (loop modes 7 and 8)."""
import numpy as np
import pytest

from oracle import tier_f
from oracle.lookahead_model import OPTIMAL, PIVOT_CAP, UNBOUNDED, LookAheadModel

STATUS = {tier_f.OPTIMAL: OPTIMAL, tier_f.UNBOUNDED: UNBOUNDED, tier_f.PIVOT_CAP: PIVOT_CAP}


def _check(A, b, c, block, cap=-1):
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy())
    status, k = ref.run(cap)
    mdl = LookAheadModel(A, b, c, block=block)
    got, kk = mdl.run(cap)          # (the model itself asserts: running vectors == last column / row of the tableau)
    assert got == STATUS[status] and kk == k
    assert mdl.log == ref.log
    assert np.array_equal(mdl.A, ref.A) and np.array_equal(mdl.b, ref.b) and np.array_equal(mdl.c, ref.c)
    assert mdl.v == ref.v[0]
    return mdl


@pytest.mark.parametrize("block", [1, 2, 5, 16])
@pytest.mark.parametrize("m,n,seed", [(5, 7, 0), (12, 9, 1), (40, 80, 3), (100, 60, 4)])
def test_lookahead_model_equals_tier_f(m, n, seed, block):
    A, b, c = tier_f.gen_dense_feasible(m, n, seed)
    mdl = _check(A, b, c, block)
    k = len(mdl.log)
    assert mdl.passes == -(-k // block)                 # one pass per block
    # the panel runs one block ahead of the pass: one more launch than passes (lps_run: ceil(P / S) + 1);
    # a block that ends exactly at the verdict needs one further launch to find it
    assert mdl.launches in (mdl.passes + 1, mdl.passes + 2)


def test_lookahead_model_cap_unbounded_and_resumed_runs():
    A, b, c = tier_f.gen_dense_feasible(30, 50, 2)
    _check(A, b, c, 7, cap=23)
    A2 = A.copy()
    A2[:, 49] = -A2[:, 49]
    _check(A2, b, c, 6)
    A3 = A.copy()
    A3[:, 0] = -A3[:, 0]
    _check(A3, b, c, 6)
    # capped runs one after the other on the same state (bench.py's steps): buffers swap roles, the
    # running vectors are rebuilt from the tableau at the start of every run
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy())
    mdl = LookAheadModel(A, b, c, block=4)
    for cap in (5, 8, 1, 16, 3):
        status, k = ref.run(cap)
        got, kk = mdl.run(cap)
        assert (got, kk) == (STATUS[status], k) and mdl.log == ref.log
        assert np.array_equal(mdl.A, ref.A) and np.array_equal(mdl.b, ref.b) and np.array_equal(mdl.c, ref.c)
    status, k = ref.run()
    got, kk = mdl.run()
    assert (got, kk) == (STATUS[status], k) and mdl.log == ref.log and mdl.v == ref.v[0]


def test_lookahead_model_degenerate_rows_and_columns_repeat_across_the_two_sets():
    k = 6
    m, n = 2 * k, k * k
    A = np.zeros((m, n))
    for i in range(k):
        for j in range(k):
            A[i, i * k + j] = 1.0
            A[k + j, i * k + j] = 1.0
    b = np.ones(m)
    c = np.random.default_rng(3).integers(1, 6, size=n).astype(np.float64)
    for block in (3, 16):
        mdl = _check(A, b, c, block, cap=2000)
        # the point of the case: a row or column pivoted on in block k is pivoted on again in block k+1,
        # i.e. while block k is still only pending for the panel that decides block k+1
        rows = [l for _, l in mdl.log]
        cols = [e for e, _ in mdl.log]
        assert any(set(rows[s:s + block]) & set(rows[s + block:s + 2 * block]) or
                   set(cols[s:s + block]) & set(cols[s + block:s + 2 * block]) for s in range(0, len(rows), block))


from hypothesis import given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402


@settings(max_examples=80, deadline=None)
@given(m=st.integers(1, 9), n=st.integers(1, 9), block=st.integers(1, 16), seed=st.integers(0, 10 ** 6),
       kind=st.sampled_from(["small_ints", "halves", "dense"]))
def test_lookahead_model_property(m, n, block, seed, kind):
    """random tiny LPs with zeros, negative entries, ties, unbounded columns and repeated pivots on the
    same row / column: the look-ahead schedule and the pivot-per-pass oracle never differ"""
    rng = np.random.default_rng(seed)
    if kind == "small_ints":
        A = rng.integers(-2, 4, size=(m, n)).astype(np.float64)
        b = rng.integers(0, 5, size=m).astype(np.float64)          # zeros in b: degenerate ties
        c = rng.integers(-2, 4, size=n).astype(np.float64)
    elif kind == "halves":
        A = rng.integers(-4, 9, size=(m, n)) / 2.0
        b = rng.integers(0, 9, size=m) / 4.0
        c = rng.integers(-4, 9, size=n) / 2.0
    else:
        A = rng.random((m, n)) - 0.2
        b = rng.random(m)
        c = rng.random(n) - 0.3
    _check(A, b, c, block, cap=200)
