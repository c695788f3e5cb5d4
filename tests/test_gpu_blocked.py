"""The blocked loop (lps_blocked.cuh): up to `block_pivots` pivots are deferred and applied in ONE
pass over the tableau.  loop_mode=5 runs the panel as two launches per pivot (kb_col, kb_row),
loop_mode=6 as one cooperative launch per block (kb_panel) followed by the pass, loop_mode=7 is the
look-ahead loop: one cooperative launch per block runs the TMA pass of block k (out of place) and the
panel of block k+1 side by side on disjoint SMs (kb_step_flush / kb_step), loop_mode=8 the same with pass warps and
panel warps inside every CTA (kb_step_ws).  Every value must still be
bit-identical to the pivot-per-pass kernels and to the binary64 oracle — pivot sequence, verdict,
every cell.  `-m gpu`."""
import threading

import numpy as np
import pytest

from oracle import tier_f

pytestmark = pytest.mark.gpu

VERDICT = {tier_f.OPTIMAL: 1, tier_f.UNBOUNDED: 2, tier_f.PIVOT_CAP: 3}
MODES = [5, 6, 7, 8]


def _L():
    import linear_programming_solver_b200 as L
    return L


def _ndev():
    import torch
    return torch.cuda.device_count()


def _same_state(st, ref):
    assert st.pivot_log == ref.log
    assert np.array_equal(st.A, ref.A) and np.array_equal(st.b, ref.b) and np.array_equal(st.c, ref.c)
    assert st.v == ref.v[0]
    assert np.array_equal(st.positions, ref.pos2var)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("block", [2, 3, 16, 32])
@pytest.mark.parametrize("m,n,seed", [(5, 7, 0), (12, 9, 1), (40, 80, 3), (100, 60, 4), (150, 150, 5),
                                      (257, 1030, 6), (300, 300, 7)])
def test_blocked_bit_exact_vs_tier_f(m, n, seed, block, mode):
    L = _L()
    A, b, c = tier_f.gen_dense_feasible(m, n, seed)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy(), nthreads=4)
    status, k = ref.run()
    st = L.LPState(A, b, c, m, n, loop_mode=mode, block_pivots=block)
    res = st.run()
    assert res.verdict == VERDICT[status] and res.npivots == k
    _same_state(st, ref)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6, 7, 10, 11, 12, 13, 14])
def test_blocked_flush_variants(variant, mode):
    """every tile shape of the pass kernel gives the same bits (wide and tall cases, capped); 0-7 are the
    cp.async kernel kb_flush (loop modes 5 and 6), 10-14 the shapes of the TMA pass (14: producer folded into a consumer warp)"""
    L = _L()
    from linear_programming_solver_b200.lp_state import LPState
    for (m, n, seed, cap) in [(700, 5001, 2, 90), (2100, 530, 3, 70)]:
        A, b, c = tier_f.gen_dense_feasible(m, n, seed)
        ref = tier_f.TierFState(A.copy(), b.copy(), c.copy(), nthreads=4)
        status, k = ref.run(cap)
        st = LPState.synthetic_dense(m, n, seed, 1000, loop_mode=mode, update_variant=variant)
        res = st.run(cap)
        assert res.verdict == VERDICT[status] and res.npivots == k
        _same_state(st, ref)


@pytest.mark.parametrize("mode", MODES)
def test_blocked_equals_pivot_per_pass_kernels(mode):
    """same handle type, block_pivots=1 (three kernels per pivot) vs the blocked loop"""
    L = _L()
    m, n = 180, 420
    A, b, c = tier_f.gen_dense_feasible(m, n, 11)
    one = L.LPState(A, b, c, m, n, loop_mode=1, block_pivots=1)
    blk = L.LPState(A, b, c, m, n, loop_mode=mode, block_pivots=7)
    r1, r2 = one.run(), blk.run()
    assert (r1.verdict, r1.npivots) == (r2.verdict, r2.npivots)
    assert one.pivot_log == blk.pivot_log
    assert np.array_equal(one.A, blk.A) and np.array_equal(one.b, blk.b) and np.array_equal(one.c, blk.c)
    assert one.v == blk.v


@pytest.mark.parametrize("mode", MODES)
def test_blocked_cap_resume_and_mixing_with_explicit_steps(mode):
    L = _L()
    A, b, c = tier_f.gen_dense_feasible(60, 60, 3)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy())
    ref.run()
    st = L.LPState(A, b, c, 60, 60, loop_mode=mode, block_pivots=4)
    r = st.run(10)                      # 2 full blocks + a partial one
    assert r.verdict == 3 and r.npivots == 10
    assert st.pivot_log == ref.log[:10]
    r = st.run(0)
    assert r.verdict == 3 and r.npivots == 0
    # host-driven step in the middle (pivot-per-pass kernels), then back to the blocked loop
    e = st.get_entering()
    l = st.get_leaving(e)
    assert (e, l) == ref.log[10]
    st.pivot(e, l)
    assert st.get_entering() == ref.log[11][0]
    r = st.run(5)
    assert r.npivots == 5 and st.pivot_log == ref.log[:16]
    r = st.run()
    assert r.verdict == 1 and r.total_pivots == len(ref.log)
    _same_state(st, ref)


@pytest.mark.parametrize("mode", MODES)
def test_blocked_immediate_and_late_verdicts(mode):
    L = _L()
    m, n = 50, 40
    A, b, c = tier_f.gen_dense_feasible(m, n, 5)
    A2 = A.copy()
    A2[:, 0] = -A2[:, 0]
    st = L.LPState(A2, b, c, m, n, loop_mode=mode)
    r = st.run()
    assert r.verdict == 2 and r.npivots == 0 and r.last_entering == 0
    st = L.LPState(A, b, -np.abs(c), m, n, loop_mode=mode)
    r = st.run()
    assert r.verdict == 1 and r.npivots == 0 and st.v == 0.0
    # unbounded only after some pivots (a column with no positive entry further right)
    A3 = A.copy()
    A3[:, n - 1] = -A3[:, n - 1]
    ref = tier_f.TierFState(A3.copy(), b.copy(), c.copy())
    status, k = ref.run()
    assert status == tier_f.UNBOUNDED and k > 0
    st = L.LPState(A3, b, c, m, n, loop_mode=mode, block_pivots=5)
    r = st.run()
    assert r.verdict == 2 and r.npivots == k
    _same_state(st, ref)


@pytest.mark.parametrize("mode", MODES)
def test_blocked_degenerate_ties_and_reentering_columns(mode):
    """exact-integer assignment-type LP: many zero ratios and ties, rows and columns that pivot
    more than once inside one block"""
    L = _L()
    k = 9
    m, n = 2 * k, k * k
    A = np.zeros((m, n))
    for i in range(k):
        for j in range(k):
            A[i, i * k + j] = 1.0
            A[k + j, i * k + j] = 1.0
    b = np.ones(m)
    rng = np.random.default_rng(3)
    c = rng.integers(1, 6, size=n).astype(np.float64)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy())
    status, npiv = ref.run(5000)
    for block in (2, 8, 32):
        st = L.LPState(A, b, c, m, n, loop_mode=mode, block_pivots=block)
        r = st.run(5000)
        assert r.verdict == VERDICT[status] and r.npivots == npiv
        _same_state(st, ref)


@pytest.mark.parametrize("mode", MODES)
def test_blocked_mid_size_capped(mode):
    L = _L()
    m = n = 1000
    A, b, c = tier_f.gen_dense_feasible(m, n, 0)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy(), nthreads=tier_f.lib().tf_max_threads())
    ref.run(400)
    st = L.LPState(A, b, c, m, n, loop_mode=mode)
    r = st.run(400)
    assert r.npivots == 400 and r.verdict == 3
    _same_state(st, ref)


def test_blocked_is_the_default_for_large_tableaus():
    """auto mode (loop_mode=0) picks the blocked loop above L2 size: one pass per 16 pivots"""
    L = _L()
    from linear_programming_solver_b200.lp_state import LPState
    m, n = 3000, 4000            # 96 MB
    A, b, c = tier_f.gen_dense_feasible(m, n, 1, nthreads=tier_f.lib().tf_max_threads())
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy(), nthreads=tier_f.lib().tf_max_threads())
    ref.run(100)
    st = LPState.synthetic_dense(m, n, 1, 1000, time_kernels=True)
    r = st.run(100)
    assert r.npivots == 100 and r.verdict == 3
    assert r.update_launches == 7          # ceil(100 / 16) passes
    _same_state(st, ref)


@pytest.mark.parametrize("mode", MODES)
def test_blocked_phase1_solver_path(mode):
    """LPSolver.solve through the blocked loop: aux LP, forced pivot, restore (LPSolver.java:116-246)"""
    L = _L()
    from oracle.arith import F64
    from oracle.simplex_ref import LPSolver as OracleSolver
    from oracle.simplex_ref import LPStandardForm as OracleForm
    rng = np.random.default_rng(7)
    m, n = 30, 20
    A = rng.integers(-4, 9, size=(m, n)).astype(np.float64)
    x = rng.integers(0, 4, size=n).astype(np.float64)
    b = A @ x + rng.integers(0, 5, size=m)
    b[::3] -= 40.0                       # some negative right-hand sides: phase 1
    c = rng.integers(-3, 6, size=n).astype(np.float64)
    want = None
    try:
        want = OracleSolver(F64).solve(OracleForm(A.tolist(), b.tolist(), c.tolist(), m, n, True))
    except Exception as ex:              # infeasible / unbounded: same exception type and text below
        want = ex
    solver = L.LPSolver(loop_mode=mode, block_pivots=6)
    form = L.LPStandardForm(A, b, c, m, n, True)
    try:
        got = solver.solve(form)
        assert not isinstance(want, Exception), want
        assert str(got) == str(want)
    except (L.LPException, L.SolutionException) as ex:
        assert isinstance(want, Exception) and str(ex) == str(want)


@pytest.mark.parametrize("mode", MODES)
def test_blocked_shard_world1(mode):
    from linear_programming_solver_b200.sharded import ShardedLPState
    m, n = 257, 1030
    A, b, c = tier_f.gen_dense_feasible(m, n, 6)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy())
    status, k = ref.run()
    st = ShardedLPState(m, n, 0, 1, A, b, c, loop_mode=mode, block_pivots=5)
    res = st.run()
    assert res.verdict == 1 and res.npivots == k
    assert st.pivot_log == ref.log
    assert np.array_equal(st.A, ref.A) and np.array_equal(st.b, ref.b) and np.array_equal(st.c, ref.c)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_blocked_multi_gpu(world, mode):
    if _ndev() < world:
        pytest.skip("needs %d GPUs" % world)
    from linear_programming_solver_b200.sharded import ShardedLPState
    m, n, seed = 403, 600, 5
    A, b, c = tier_f.gen_dense_feasible(m, n, seed)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy(), nthreads=4)
    status, k = ref.run()
    shards = [ShardedLPState(m, n, r, world, synthetic_seed=seed, device=r, loop_mode=mode, block_pivots=6)
              for r in range(world)]
    ptrs = [s.comm_ptr() for s in shards]
    for s in shards:
        s.attach_ptrs(ptrs)
    results = [None] * world

    def work(r):
        results[r] = shards[r].run()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    for r, s in enumerate(shards):
        assert results[r] is not None and results[r].verdict == 1 and results[r].npivots == k
        assert s.pivot_log == ref.log
        assert np.array_equal(s.A, ref.A[s.row0:s.row1])
        assert np.array_equal(s.b, ref.b[s.row0:s.row1])
        assert np.array_equal(s.c, ref.c) and s.v == ref.v[0]


@pytest.mark.parametrize("mode", [6, 7, 8])
@pytest.mark.parametrize("ctas,chunk", [(1, 12), (3, 24), (40, 0), (147, 48)])
def test_look_ahead_role_split_and_chunking(mode, ctas, chunk):
    """the result may not depend on how many CTAs run the panel, nor on the chunk height of the pass"""
    from linear_programming_solver_b200.lp_state import LPState
    m, n, seed, cap = 900, 2100, 4, 150
    A, b, c = tier_f.gen_dense_feasible(m, n, seed)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy(), nthreads=4)
    status, k = ref.run(cap)
    st = LPState.synthetic_dense(m, n, seed, 1000, loop_mode=mode, panel_ctas=ctas, pass_chunk_rows=chunk)
    res = st.run(cap)
    assert res.verdict == VERDICT[status] and res.npivots == k
    _same_state(st, ref)


@pytest.mark.parametrize("mode", [6, 7, 8])
def test_look_ahead_repeated_runs_and_buffer_swaps(mode):
    """many short runs: the tableau ends up in either buffer of the out-of-place pass, partial blocks,
    caps at every position of a block"""
    L = _L()
    m, n = 120, 200
    A, b, c = tier_f.gen_dense_feasible(m, n, 9)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy())
    ref.run()
    st = L.LPState(A, b, c, m, n, loop_mode=mode, block_pivots=4)
    done = 0
    for step in [1, 2, 3, 4, 5, 7, 8, 9, 1, 16, 17]:
        r = st.run(step)
        done += r.npivots
        assert st.pivot_log == ref.log[:done]
        if r.verdict != 3:
            break
        again = tier_f.TierFState(A.copy(), b.copy(), c.copy())
        again.run(done)
        assert np.array_equal(st.A, again.A) and np.array_equal(st.b, again.b) and np.array_equal(st.c, again.c)
    r = st.run()
    assert r.verdict == 1 and r.total_pivots == len(ref.log)
    _same_state(st, ref)


@pytest.mark.parametrize("mode", [6, 7, 8])
def test_handle_reloaded_with_a_larger_lp(mode):
    """one handle, a small LP and then a large one: the pass kernel chosen for the second size must get
    its shared-memory opt-in too"""
    L = _L()
    from linear_programming_solver_b200.lp_state import _dp
    A, b, c = tier_f.gen_dense_feasible(40, 60, 1)
    st = L.LPState(A, b, c, 40, 60, loop_mode=mode, update_variant=0 if mode == 6 else -1)
    st.run(20)
    m, n = 2600, 2100
    A, b, c = tier_f.gen_dense_feasible(m, n, 2)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy(), nthreads=4)
    ref.run(40)
    st._ck(st._lib.lps_load(st._h, m, n, _dp(A), n, _dp(b), _dp(c), 0.0), "lps_load")
    r = st.run(40)
    assert r.npivots == 40
    _same_state(st, ref)
