"""Worker for the multi-process sharding test: run under torchrun with one rank per GPU.
Each rank builds its row shard, attaches peers over CUDA IPC, runs the pivot loop and checks its
rows, the (global) pivot log, positions and objective against the binary64 oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from linear_programming_solver_b200.sharded import ShardedLPState  # noqa: E402
from oracle import tier_f  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("gloo")
    for (m, n, seed, cap, synthetic) in [(37, 50, 0, -1, False), (200, 300, 1, -1, True), (1000, 1500, 2, 300, True),
                                         (64, 40, 3, -1, False)]:
        A, b, c = tier_f.gen_dense_feasible(m, n, seed)
        ref = tier_f.TierFState(A.copy(), b.copy(), c.copy(), nthreads=4)
        status, k = ref.run(cap)
        if synthetic:
            st = ShardedLPState(m, n, rank, world, synthetic_seed=seed, device=local_rank)
        else:
            r0, r1 = (rank * m) // world, ((rank + 1) * m) // world
            st = ShardedLPState(m, n, rank, world, A[r0:r1], b[r0:r1], c, device=local_rank)
        st.attach_via(dist)
        res = st.run(cap)
        want = {tier_f.OPTIMAL: 1, tier_f.UNBOUNDED: 2, tier_f.PIVOT_CAP: 3}[status]
        assert res.verdict == want, (rank, res.verdict, want)
        assert res.npivots == k, (rank, res.npivots, k)
        assert st.pivot_log == ref.log, "rank %d: pivot log differs" % rank
        assert np.array_equal(st.A, ref.A[st.row0:st.row1]), "rank %d: A rows differ" % rank
        assert np.array_equal(st.b, ref.b[st.row0:st.row1])
        assert np.array_equal(st.c, ref.c)
        assert st.v == ref.v[0]
        assert np.array_equal(st.positions, ref.pos2var)
        x = st.primal(n, dist)
        lookup = {int(v): p for p, v in enumerate(ref.pos2var)}
        xr = np.array([ref.b[lookup[j] - n] if lookup[j] >= n else 0.0 for j in range(n)])
        assert np.array_equal(x, xr)
        st.close()
        dist.barrier()
    if rank == 0:
        print("SHARD_WORKER_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
