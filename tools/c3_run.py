"""BASELINE.json configs[2]: synthetic dense LP 10,000 x 10,000 with mixed <= / >= / == rows (the '>='
rows have b < 0, so phase 1 — the auxiliary LP of LPSolver.java:283-321 — is forced), on one B200.

The first-positive entering rule needs a very large number of pivots on this instance, so the run is
pivot-capped: it times phase 1 on the device (aux tableau built in HBM, forced first pivot
LPSolver.java:138, then the loop :141-161) and checks the first `--check` pivots of that loop against
the binary64 CPU twin pivot for pivot (the twin moves 1.6 GB per pivot through host memory, so only a
prefix is affordable).  With --full it keeps going until a verdict (or --max-pivots).

    python tools/c3_run.py [--rows 10000 --cols 10000 --pivots 200000 --check 256] [--out file.json]
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import linear_programming_solver_b200 as L  # noqa: E402
from oracle import tier_f  # noqa: E402


def digest(log):
    return hashlib.sha256(np.asarray(log, dtype=np.int32).tobytes()).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10000)
    ap.add_argument("--cols", type=int, default=10000)
    ap.add_argument("--pivots", type=int, default=200000)
    ap.add_argument("--check", type=int, default=256)
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--max-pivots", type=int, default=4000000)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    m, n = a.rows, a.cols
    threads = tier_f.lib().tf_max_threads()
    A, b, c = tier_f.gen_mixed_rows(m, n, 0, True)
    k = tier_f.min_in_b(b)
    rec = {"m": m, "n": n, "negative_rhs_rows": int((b < 0).sum()), "min_in_b": int(k), "host_threads": threads}
    assert k >= 0 and b[k] < 0, "phase 1 is not forced on this instance"

    # ---- device: convertIntoAuxLP + forced pivot + the loop, capped ----
    t0 = time.perf_counter()
    st = L.LPState.aux(A, b, m, n, time_kernels=True)            # m x (n+1) aux tableau, x0 = column n
    load_s = time.perf_counter() - t0
    st.pivot(n, k)                                               # LPSolver.java:138
    r = st.run(a.pivots)
    log = st.pivot_log
    rec["gpu_phase1"] = {"verdict": int(r.verdict), "pivots": int(r.npivots) + 1, "device_ms": r.device_ms,
                         "pivots_per_s": r.npivots / (r.device_ms / 1e3), "passes": int(r.update_launches),
                         "pass_ms": r.update_ms / max(r.update_launches, 1), "load_s": load_s,
                         "aux_objective": st.v, "log_sha256_first": digest(log[:a.check])}
    if a.full and r.verdict == 3:
        t0 = time.perf_counter()
        r2 = st.run(a.max_pivots - a.pivots)
        rec["gpu_phase1_full"] = {"verdict": int(r2.verdict), "pivots": int(r2.total_pivots), "seconds": time.perf_counter() - t0,
                                  "aux_objective": st.v}
    st.close()

    # ---- CPU twin, same prefix ----
    auxA = np.empty((m, n + 1))
    auxA[:, :n] = A
    auxA[:, n] = -1.0
    auxc = np.zeros(n + 1)
    auxc[n] = -1.0
    t0 = time.perf_counter()
    ref = tier_f.TierFState(auxA, b.copy(), auxc, nthreads=threads)
    ref.pivot(n, k)
    ref.run(a.check - 1)
    rec["cpu_twin"] = {"pivots": len(ref.log), "seconds": time.perf_counter() - t0,
                       "pivots_per_s": len(ref.log) / (time.perf_counter() - t0), "log_sha256_first": digest(ref.log)}
    rec["same_first_pivots"] = rec["cpu_twin"]["log_sha256_first"] == rec["gpu_phase1"]["log_sha256_first"]
    print(json.dumps(rec, indent=1))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(rec, f, indent=1)


if __name__ == "__main__":
    main()
