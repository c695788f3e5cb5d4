"""The persistent cooperative loop kernel (loop_mode=2): same results, bit for bit, as the
three-kernel path and the binary64 oracle.  `-m gpu`."""
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


@pytest.mark.parametrize("m,n,seed", [(5, 7, 0), (12, 9, 1), (40, 80, 3), (100, 60, 4), (150, 150, 5),
                                      (257, 1030, 6), (300, 300, 7), (700, 5001, 2)])
def test_persistent_bit_exact_vs_tier_f(m, n, seed):
    L = _L()
    A, b, c = tier_f.gen_dense_feasible(m, n, seed)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy(), nthreads=4)
    cap = 80 if m >= 700 else -1
    status, k = ref.run(cap)
    st = L.LPState(A, b, c, m, n, loop_mode=2)
    res = st.run(cap)
    assert res.verdict == {tier_f.OPTIMAL: 1, tier_f.UNBOUNDED: 2, tier_f.PIVOT_CAP: 3}[status]
    assert res.npivots == k
    assert st.pivot_log == ref.log
    assert np.array_equal(st.A, ref.A) and np.array_equal(st.b, ref.b) and np.array_equal(st.c, ref.c)
    assert st.v == ref.v[0]
    assert np.array_equal(st.positions, ref.pos2var)


def test_persistent_cap_resume_and_mixing_with_explicit_steps():
    L = _L()
    A, b, c = tier_f.gen_dense_feasible(60, 60, 3)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy())
    ref.run()
    st = L.LPState(A, b, c, 60, 60, loop_mode=2)
    r = st.run(10)
    assert r.verdict == 3 and r.npivots == 10
    assert st.pivot_log == ref.log[:10]
    r = st.run(0)
    assert r.verdict == 3 and r.npivots == 0
    # host-driven step in the middle (three-kernel path), then back to the persistent loop
    e = st.get_entering()
    l = st.get_leaving(e)
    assert (e, l) == ref.log[10]
    st.pivot(e, l)
    assert st.get_entering() == ref.log[11][0]
    r = st.run(5)
    assert r.npivots == 5 and st.pivot_log == ref.log[:16]
    r = st.run()
    assert r.verdict == 1 and r.total_pivots == len(ref.log)
    assert st.pivot_log == ref.log
    assert np.array_equal(st.A, ref.A) and st.v == ref.v[0]


def test_persistent_immediate_verdicts():
    L = _L()
    m, n = 50, 40
    A, b, c = tier_f.gen_dense_feasible(m, n, 5)
    A2 = A.copy()
    A2[:, 0] = -A2[:, 0]
    st = L.LPState(A2, b, c, m, n, loop_mode=2)
    r = st.run()
    assert r.verdict == 2 and r.npivots == 0 and r.last_entering == 0
    st = L.LPState(A, b, -np.abs(c), m, n, loop_mode=2)
    r = st.run()
    assert r.verdict == 1 and r.npivots == 0 and st.v == 0.0


def test_persistent_mid_size_capped():
    L = _L()
    m = n = 1000
    A, b, c = tier_f.gen_dense_feasible(m, n, 0)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy(), nthreads=tier_f.lib().tf_max_threads())
    ref.run(400)
    st = L.LPState(A, b, c, m, n, loop_mode=2)
    r = st.run(400)
    assert r.npivots == 400 and r.verdict == 3
    assert st.pivot_log == ref.log
    assert np.array_equal(st.A, ref.A) and np.array_equal(st.b, ref.b) and np.array_equal(st.c, ref.c)


def test_persistent_shard_world1():
    from linear_programming_solver_b200.sharded import ShardedLPState
    m, n = 257, 1030
    A, b, c = tier_f.gen_dense_feasible(m, n, 6)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy())
    status, k = ref.run()
    st = ShardedLPState(m, n, 0, 1, A, b, c, loop_mode=2)
    res = st.run()
    assert res.verdict == 1 and res.npivots == k
    assert st.pivot_log == ref.log
    assert np.array_equal(st.A, ref.A) and np.array_equal(st.b, ref.b) and np.array_equal(st.c, ref.c)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_persistent_multi_gpu(world):
    if _ndev() < world:
        pytest.skip("needs %d GPUs" % world)
    from linear_programming_solver_b200.sharded import ShardedLPState
    m, n, seed = 403, 600, 5
    A, b, c = tier_f.gen_dense_feasible(m, n, seed)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy(), nthreads=4)
    status, k = ref.run()
    shards = [ShardedLPState(m, n, r, world, synthetic_seed=seed, device=r, loop_mode=2) for r in range(world)]
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
