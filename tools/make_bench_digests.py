"""Digests of the bench workload's pivot sequence (BASELINE configs[3]: 20,000 x 40,000, seed 0, every cost
positive): SHA-256 of the (entering, leaving) log and of the b column after k pivots, k = 256, 512, ...

    python tools/make_bench_digests.py cpu  [max_pivots] [out.json]    binary64 CPU twin (oracle/tier_f.c), no GPU
    python tools/make_bench_digests.py gpu  [max_pivots] [out.json]    pivot-per-pass kernels (loop_mode=1), one GPU

bench.py compares the digests of ITS run (blocked / look-ahead loop, 1-8 GPUs) with the committed table
tests/golden/bench_c4_seed0_digests.json and exits non-zero on a mismatch: the partition and the loop shape
must not change results (LPState.java:222-223)."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STEP = 256


def digest_log(pairs) -> str:
    return hashlib.sha256(np.asarray(pairs, dtype=np.int32).reshape(-1, 2).tobytes()).hexdigest()


def digest_b(b) -> str:
    return hashlib.sha256(np.ascontiguousarray(b, dtype=np.float64).tobytes()).hexdigest()


def main():
    backend = sys.argv[1]
    cap = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "bench_digests_%s.json" % backend)
    m, n, seed = 20000, 40000, 0
    table, t0 = {}, time.perf_counter()
    if backend == "cpu":
        from oracle import tier_f
        threads = len(os.sched_getaffinity(0))
        A, b, c = tier_f.gen_dense_feasible(m, n, seed, nthreads=threads)
        st = tier_f.TierFState(A, b, c, nthreads=threads)
        done = 0
        while done < cap:
            status, k = st.run(STEP)
            done += k
            if k < STEP:
                break
            table[str(done)] = {"log_sha256": digest_log(st.log[:done]), "b_sha256": digest_b(st.b)}
            json.dump({"backend": backend, "m": m, "n": n, "seed": seed, "digests": table,
                       "wall_s": time.perf_counter() - t0}, open(out, "w"), indent=0)
    else:
        import linear_programming_solver_b200 as L
        st = L.LPState.synthetic_dense(m, n, seed, 1000, loop_mode=1, block_pivots=1)
        done = 0
        while done < cap:
            r = st.run(STEP)
            done += r.npivots
            if r.npivots < STEP:
                break
            table[str(done)] = {"log_sha256": digest_log(st.pivot_log), "b_sha256": digest_b(st.b)}
        json.dump({"backend": backend, "m": m, "n": n, "seed": seed, "digests": table,
                   "wall_s": time.perf_counter() - t0}, open(out, "w"), indent=0)
    print(json.dumps({"backend": backend, "pivots": done, "checkpoints": len(table), "wall_s": time.perf_counter() - t0}))


if __name__ == "__main__":
    main()
