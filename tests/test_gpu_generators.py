"""The synthetic LP families of SURVEY.md §8d generated in HBM (lps_generate_lp): the device
generator must equal its numpy restatement cell for cell, and the solves on them must agree with the
oracle — unbounded verdicts (immediately and after a long run), degenerate exact-integer pivoting
identical to the reference's decimal arithmetic.  `-m gpu`."""
import threading

import numpy as np
import pytest

from oracle import tier_f

pytestmark = pytest.mark.gpu


def _L():
    import linear_programming_solver_b200 as L
    return L


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("m,n", [(40, 70), (301, 517)])
def test_generators_match_their_restatements(m, n):
    from linear_programming_solver_b200 import _native as N
    from linear_programming_solver_b200.lp_state import LPState
    for kind, param, ref in [(N.LPS_GEN_DENSE, 300, tier_f.gen_dense_feasible(m, n, 5, 300)),
                             (N.LPS_GEN_UNBOUNDED, n - 1, tier_f.gen_unbounded(m, n, 5, n - 1)),
                             (N.LPS_GEN_UNBOUNDED, 0, tier_f.gen_unbounded(m, n, 5, 0)),
                             (N.LPS_GEN_ASSIGNMENT, 0, tier_f.gen_assignment(m, n))]:
        st = LPState.synthetic(kind, m, n, 5, param)
        A, b, c = ref
        assert np.array_equal(st.A, A) and np.array_equal(st.b, b) and np.array_equal(st.c, c)
        assert st.v == 0.0


@pytest.mark.parametrize("mode", [1, 6, 7, 8])
@pytest.mark.parametrize("col", ["first", "last"])
def test_unbounded_family(col, mode):
    from linear_programming_solver_b200 import _native as N
    from linear_programming_solver_b200.lp_state import LPState
    m, n = 120, 200
    j = 0 if col == "first" else n - 1
    A, b, c = tier_f.gen_unbounded(m, n, 2, j)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy())
    status, k = ref.run()
    assert status == tier_f.UNBOUNDED and (k == 0) == (col == "first")
    st = LPState.synthetic(N.LPS_GEN_UNBOUNDED, m, n, 2, j, loop_mode=mode)
    r = st.run()
    assert r.verdict == 2 and r.npivots == k
    assert st.pivot_log == ref.log
    assert np.array_equal(st.A, ref.A) and np.array_equal(st.b, ref.b) and np.array_equal(st.c, ref.c)
    # through the reference-facing API: SolutionException with the reference's text (LPSolver.java:105)
    L = _L()
    with pytest.raises(L.SolutionException, match="This linear program is unbounded"):
        L.LPSolver().solve(L.LPStandardForm(A, b, c, m, n, True))


@pytest.mark.parametrize("mode", [1, 2, 6, 7, 8])
def test_assignment_family_is_exact_and_matches_the_decimal_oracle(mode):
    """degenerate, totally unimodular: entries stay in {-1,0,1}; pivot sequence identical to the
    reference's 15-digit decimal arithmetic (Tier D) as well as to the binary64 twin"""
    from linear_programming_solver_b200 import _native as N
    from linear_programming_solver_b200.lp_state import LPState
    from oracle import tier_d
    m, n = 60, 150
    A, b, c = tier_f.gen_assignment(m, n)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy())
    status, k = ref.run(20000)
    assert status == tier_f.OPTIMAL
    st = LPState.synthetic(N.LPS_GEN_ASSIGNMENT, m, n, 0, 0, loop_mode=mode, block_pivots=7)
    r = st.run(20000)
    assert r.verdict == 1 and r.npivots == k
    assert st.pivot_log == ref.log
    Af = st.A
    assert np.array_equal(Af, ref.A) and np.array_equal(st.b, ref.b) and np.array_equal(st.c, ref.c)
    assert set(np.unique(Af)).issubset({-1.0, 0.0, 1.0})
    assert st.v == ref.v[0] == float(int(st.v))
    # zero-ratio ties really occur on this family
    dec = tier_d.TierDState(A, b, c)
    dec.run(20000)
    assert dec.log == ref.log and float(dec.v) == st.v


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("kind", ["unbounded", "assignment"])
def test_families_row_sharded(kind, world):
    if _ndev() < world:
        pytest.skip("needs %d GPUs" % world)
    from linear_programming_solver_b200 import _native as N
    from linear_programming_solver_b200.sharded import ShardedLPState
    m, n = 203, 320
    if kind == "unbounded":
        A, b, c = tier_f.gen_unbounded(m, n, 4, n - 1)
        k_, param, want = N.LPS_GEN_UNBOUNDED, n - 1, 2
    else:
        A, b, c = tier_f.gen_assignment(m, n)
        k_, param, want = N.LPS_GEN_ASSIGNMENT, 0, 1
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy())
    status, k = ref.run(50000)
    shards = [ShardedLPState(m, n, r, world, synthetic_seed=4, pos_permille=param, synthetic_kind=k_, device=r)
              for r in range(world)]
    ptrs = [s.comm_ptr() for s in shards]
    for s in shards:
        s.attach_ptrs(ptrs)
    results = [None] * world

    def work(r):
        results[r] = shards[r].run(50000)

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    for r, s in enumerate(shards):
        assert results[r] is not None and results[r].verdict == want and results[r].npivots == k
        assert s.pivot_log == ref.log
        assert np.array_equal(s.A, ref.A[s.row0:s.row1]) and np.array_equal(s.b, ref.b[s.row0:s.row1])
        assert np.array_equal(s.c, ref.c) and s.v == ref.v[0]
