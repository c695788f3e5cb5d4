"""Host-side sharding logic on CPU with world_size-2/3 gloo process groups: the row partition, the
candidate reduction and the owner broadcast give the same pivots as the unsharded oracle."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from linear_programming_solver_b200.sharded import owner_of, partition, reduce_candidates
from oracle import tier_f


def test_partition_matches_reference_block_split():
    for m in (0, 1, 7, 14, 1981, 4160, 20000):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for k in range(world):
                lo, hi = partition(m, world, k)
                assert lo == (k * m) // world and hi == ((k + 1) * m) // world   # LPState.java:222-223
                covered.extend(range(lo, hi))
            assert covered == list(range(m))
            for row in (0, m // 2, m - 1):
                if m:
                    lo, hi = partition(m, world, owner_of(row, m, world))
                    assert lo <= row < hi


def test_reduce_candidates_tie_break():
    assert reduce_candidates([(2.0, 5), (1.0, 9), (1.0, 3)]) == (1.0, 3)      # lowest row wins ties
    assert reduce_candidates([(float("inf"), -1), (float("inf"), -1)]) == (float("inf"), -1)
    assert reduce_candidates([(-1.0, 8), (0.0, 0)]) == (-1.0, 8)               # negative ratios win
    assert reduce_candidates([(3.0, -1), (4.0, 2)]) == (4.0, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, m, n, seed, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A, b, c = tier_f.gen_dense_feasible(m, n, seed)
    lo, hi = partition(m, world, rank)
    Al, bl, cl, v = A[lo:hi].copy(), b[lo:hi].copy(), c.copy(), 0.0
    eps, inf = 1e-9, 1e50
    log = []
    while True:
        e = tier_f.lib().tf_get_entering(tier_f._dp(cl), n, eps)
        if e == -1:
            break
        # local ratio test -> (ratio, GLOBAL row)
        col = Al[:, e]
        ratio, row = inf, -1
        for i in range(hi - lo):
            if not (col[i] < eps):
                s = bl[i] / col[i]
                if s < ratio:
                    ratio, row = s, lo + i
        cands = [None] * world
        dist.all_gather_object(cands, (float(ratio), int(row)))
        _, l = reduce_candidates(cands)
        if l < 0:
            break
        owner = owner_of(l, m, world)
        payload = [None]
        if rank == owner:
            p = Al[l - lo, e]
            r = Al[l - lo] / p
            r[e] = 1.0 / p
            bl[l - lo] = bl[l - lo] / p
            Al[l - lo] = r
            payload = [(r.copy(), float(bl[l - lo]), float(p))]
        dist.broadcast_object_list(payload, src=owner)
        r, b_l, p = payload[0]
        for i in range(hi - lo):
            if lo + i == l:
                continue
            a = Al[i, e]
            Al[i] = Al[i] - a * r
            Al[i, e] = -(a / p)
            bl[i] = bl[i] - a * b_l
        ce = cl[e]
        v = v + b_l * ce
        cl = cl - ce * r
        cl[e] = -(ce / p)
        log.append((e, l))
    out_q.put((rank, log, Al, bl, cl, v))
    dist.barrier()
    dist.destroy_process_group()


def _blocked_worker(rank, world, port, m, n, seed, block, out_q):
    """the row-sharded BLOCKED loop (oracle/blocked_model.py) with gloo carrying the two exchanges"""
    from oracle.blocked_model import ShardedBlockedModel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A, b, c = tier_f.gen_dense_feasible(m, n, seed)
    lo, hi = partition(m, world, rank)

    def all_gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    def broadcast(obj, src):
        box = [obj]
        dist.broadcast_object_list(box, src=src)
        return box[0]

    mdl = ShardedBlockedModel(A[lo:hi], b[lo:hi], c, lo, hi, m, lambda row: owner_of(row, m, world), all_gather,
                              broadcast, rank, block=block)
    status, k = mdl.run()
    out_q.put((rank, mdl.log, mdl.A.copy(), mdl.b.copy(), mdl.c.copy(), mdl.v, mdl.passes))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,m,n,seed,block", [(2, 14, 18, 0, 4), (2, 31, 20, 1, 16), (3, 25, 40, 2, 5)])
def test_sharded_blocked_loop_equals_unsharded_oracle(world, m, n, seed, block):
    """what the product does on N GPUs — rows sharded, 16 pivots per pass, candidates all-gathered, the
    owner replaying + scaling + broadcasting the leaving row — restated on the CPU over gloo: same pivots
    and same cells as the unsharded binary64 oracle"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_blocked_worker, args=(r, world, port, m, n, seed, block, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    A, b, c = tier_f.gen_dense_feasible(m, n, seed)
    ref = tier_f.TierFState(A, b, c)
    ref.run()
    for rank, log, Al, bl, cl, v, passes in results:
        lo, hi = partition(m, world, rank)
        assert log == ref.log
        assert np.array_equal(Al, ref.A[lo:hi]) and np.array_equal(bl, ref.b[lo:hi])
        assert np.array_equal(cl, ref.c) and v == ref.v[0]
        assert passes == -(-len(ref.log) // block)


def _lookahead_worker(rank, world, port, m, n, seed, block, out_q):
    """the row-sharded LOOK-AHEAD loop (oracle/lookahead_model.py) with gloo carrying the two exchanges"""
    from oracle.lookahead_model import ShardedLookAheadModel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A, b, c = tier_f.gen_dense_feasible(m, n, seed)
    lo, hi = partition(m, world, rank)

    def all_gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    def broadcast(obj, src):
        box = [obj]
        dist.broadcast_object_list(box, src=src)
        return box[0]

    mdl = ShardedLookAheadModel(A[lo:hi], b[lo:hi], c, lo, hi, m, lambda row: owner_of(row, m, world), all_gather,
                                broadcast, rank, block=block)
    mdl.run(7)                         # a capped run first: the next one starts from the other tableau buffer
    status, k = mdl.run()
    out_q.put((rank, mdl.log, mdl.A.copy(), mdl.b.copy(), mdl.c.copy(), mdl.v, mdl.passes, mdl.launches))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,m,n,seed,block", [(2, 14, 18, 0, 4), (2, 31, 20, 1, 16), (3, 25, 40, 2, 5)])
def test_sharded_lookahead_loop_equals_unsharded_oracle(world, m, n, seed, block):
    """the look-ahead loop on N GPUs (kb_step<true>: every rank's panel replays [previous block, own block] on
    its rows of the not-yet-updated tableau, local running b, replicated running c, candidates all-gathered,
    the owner replaying + scaling + broadcasting the leaving row) restated on the CPU over gloo: same pivots
    and same cells as the unsharded binary64 oracle"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_lookahead_worker, args=(r, world, port, m, n, seed, block, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    A, b, c = tier_f.gen_dense_feasible(m, n, seed)
    ref = tier_f.TierFState(A, b, c)
    ref.run()
    for rank, log, Al, bl, cl, v, passes, launches in results:
        lo, hi = partition(m, world, rank)
        assert log == ref.log
        assert np.array_equal(Al, ref.A[lo:hi]) and np.array_equal(bl, ref.b[lo:hi])
        assert np.array_equal(cl, ref.c) and v == ref.v[0]
        assert launches > passes


@pytest.mark.parametrize("world,m,n,seed", [(2, 14, 18, 0), (2, 31, 20, 1), (3, 25, 40, 2)])
def test_sharded_loop_equals_unsharded_oracle(world, m, n, seed):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, m, n, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    A, b, c = tier_f.gen_dense_feasible(m, n, seed)
    ref = tier_f.TierFState(A, b, c)
    ref.run()
    for rank, log, Al, bl, cl, v in results:
        lo, hi = partition(m, world, rank)
        assert log == ref.log
        assert np.array_equal(Al, ref.A[lo:hi]) and np.array_equal(bl, ref.b[lo:hi])
        assert np.array_equal(cl, ref.c) and v == ref.v[0]
