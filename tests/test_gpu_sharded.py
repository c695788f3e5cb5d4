"""Row-sharded pivot loop on GPUs.  `-m gpu`; the multi-rank cases need >= 2 devices."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from oracle import tier_f

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("m,n,seed", [(5, 7, 0), (40, 80, 3), (257, 1030, 6), (300, 300, 7)])
def test_world1_shard_equals_tier_f(m, n, seed):
    """The sharded kernels (mailbox, flags, fused scale+broadcast) with a single rank: bit-exact."""
    from linear_programming_solver_b200.sharded import ShardedLPState
    A, b, c = tier_f.gen_dense_feasible(m, n, seed)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy())
    status, k = ref.run()
    st = ShardedLPState(m, n, 0, 1, A, b, c)
    res = st.run()
    assert res.verdict == 1 and res.npivots == k
    assert st.pivot_log == ref.log
    assert np.array_equal(st.A, ref.A) and np.array_equal(st.b, ref.b) and np.array_equal(st.c, ref.c)
    assert st.v == ref.v[0]
    assert np.array_equal(st.positions, ref.pos2var)


def test_world1_synthetic_and_cap():
    from linear_programming_solver_b200.sharded import ShardedLPState
    m, n = 500, 700
    A, b, c = tier_f.gen_dense_feasible(m, n, 4)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy(), nthreads=4)
    ref.run(120)
    st = ShardedLPState(m, n, 0, 1, synthetic_seed=4)
    r = st.run(50)
    assert r.verdict == 3 and r.npivots == 50
    r = st.run(70)
    assert r.verdict == 3 and r.total_pivots == 120
    assert st.pivot_log == ref.log
    assert np.array_equal(st.A, ref.A)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_single_process_multi_gpu(world):
    """One process driving `world` GPUs (the JNI host's shape): peer pointers, one thread per rank."""
    if _ndev() < world:
        pytest.skip("needs %d GPUs" % world)
    from linear_programming_solver_b200.sharded import ShardedLPState
    m, n, seed = 403, 600, 5
    A, b, c = tier_f.gen_dense_feasible(m, n, seed)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy(), nthreads=4)
    status, k = ref.run()
    shards = [ShardedLPState(m, n, r, world, synthetic_seed=seed, device=r) for r in range(world)]
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
        assert np.array_equal(s.c, ref.c)
        assert s.v == ref.v[0]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_process_ipc(world):
    """One process per GPU under torchrun, CUDA IPC attach (bench.py's N > 1 shape)."""
    if _ndev() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29500 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mp_shard_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "SHARD_WORKER_OK world=%d" % world in out.stdout
