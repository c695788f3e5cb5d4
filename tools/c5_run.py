"""BASELINE.json configs[4]: 60,000 x 120,000 LPs (57.6 GB tableau) row-sharded over the GPUs of one
box — the generic dense instance (capped), the degenerate assignment instance (full solve) and the
unbounded instances (immediate, and after a long run).  Run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/c5_run.py [--rows 60000 --cols 120000]

At this size nothing on the CPU can check the cells, so the checks are size-independent properties:
  * the pivot log of the sharded run equals, pivot for pivot, the log of the same LP on ONE GPU
    (the tableau still fits a single B200's 180 GB) — results do not depend on the GPU count;
  * the assignment LP (totally unimodular) ends optimal with the objective equal to the size of a
    perfect matching, an exactly integral b column and positions that are a permutation;
  * the unbounded LPs end UNBOUNDED — after 0 pivots when the bad column is column 0.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from linear_programming_solver_b200 import _native as N  # noqa: E402
from linear_programming_solver_b200.lp_state import LPState  # noqa: E402
from linear_programming_solver_b200.sharded import ShardedLPState  # noqa: E402


def log_hash(log):
    return hashlib.sha256(np.asarray(log, dtype=np.int32).tobytes()).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=60000)
    ap.add_argument("--cols", type=int, default=120000)
    ap.add_argument("--dense-pivots", type=int, default=2048)
    ap.add_argument("--check-pivots", type=int, default=256)
    ap.add_argument("--long-cap", type=int, default=40000)
    ap.add_argument("--out", default="gpurun_out/c5_run.json")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m, n = a.rows, a.cols
    out = {"m": m, "n": n, "world": world, "tableau_gb": 8.0 * (m + 1) * (n + 1) / 1e9, "cases": {}}

    def sharded(kind, param, seed=0):
        st = ShardedLPState(m, n, rank, world, synthetic_seed=seed, pos_permille=param, synthetic_kind=kind,
                            device=local, time_kernels=True)
        st.attach_via(dist)
        return st

    def timed_run(st, cap):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        r = st.run(cap)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt, r.device_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return r, float(tt[0]), float(tt[1])

    # ---- (i) generic dense, capped ----
    st = sharded(N.LPS_GEN_DENSE, 1000)
    st.run(64)                                             # warm-up
    r, wall, dev_ms = timed_run(st, a.dense_pivots)
    log = st.pivot_log
    bytes_pp = 16.0 * (m + 1) * (n + 1)
    case = {"verdict": int(r.verdict), "pivots": int(r.npivots), "pivots_per_s": r.npivots / (dev_ms / 1e3),
            "pass_ms": r.update_ms / max(r.update_launches, 1), "passes": int(r.update_launches),
            "pivot_equivalent_gbs": bytes_pp * r.npivots / (dev_ms * 1e-3) / 1e9,
            "dram_gbs_per_gpu": (bytes_pp / world) * r.update_launches / (dev_ms * 1e-3) / 1e9,
            "log_sha256_first": log_hash(log[:64 + a.check_pivots])}
    st.close()
    dist.barrier()
    if rank == 0:                                           # the same LP on ONE GPU: same pivots?
        one = LPState.synthetic(N.LPS_GEN_DENSE, m, n, 0, 1000, device=local, time_kernels=True)
        r1 = one.run(64 + a.check_pivots)
        case["single_gpu_log_sha256_first"] = log_hash(one.pivot_log)
        case["same_pivots_as_one_gpu"] = case["single_gpu_log_sha256_first"] == case["log_sha256_first"]
        case["single_gpu_pivots_per_s"] = r1.npivots / (r1.device_ms / 1e3)
        one.close()
        out["cases"]["dense_capped"] = case
    dist.barrier()

    # ---- (ii) degenerate: bipartite assignment, full solve ----
    st = sharded(N.LPS_GEN_ASSIGNMENT, 0)
    r, wall, dev_ms = timed_run(st, 4 * m)
    b = st.gather_b(dist)
    pos = st.positions
    if rank == 0:
        out["cases"]["assignment_degenerate"] = {
            "verdict": int(r.verdict), "pivots": int(r.npivots), "seconds": dev_ms / 1e3,
            "pivots_per_s": r.npivots / (dev_ms / 1e3), "objective": st.v, "expected_objective": float(m // 2),
            "b_is_integral": bool(np.all(b == np.round(b))), "b_in_0_1": bool(np.all((b == 0.0) | (b == 1.0))),
            "positions_are_a_permutation": bool(np.array_equal(np.sort(pos), np.arange(m + n)))}
    st.close()
    dist.barrier()

    # ---- (iii) unbounded: bad column first (immediate) and last (after a long run, capped) ----
    for name, col, cap in (("unbounded_col0", 0, 16), ("unbounded_last_col", n - 1, a.long_cap)):
        st = sharded(N.LPS_GEN_UNBOUNDED, col)
        r, wall, dev_ms = timed_run(st, cap)
        if rank == 0:
            out["cases"][name] = {"verdict": int(r.verdict), "pivots": int(r.npivots), "cap": cap,
                                  "last_entering": int(r.last_entering), "seconds": dev_ms / 1e3,
                                  "pivots_per_s": r.npivots / (dev_ms / 1e3) if r.npivots else None}
        st.close()
        dist.barrier()
    if rank == 0:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
