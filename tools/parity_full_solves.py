"""One-off large parity demonstrations on a GPU box; writes a JSON record for profiles/.

  c2   1,000 x 1,000 dense LP (BASELINE config 1) solved to optimality on the GPU and by the
       decimal-15 C oracle (the reference's arithmetic): full pivot-sequence comparison.
  c4   20,000 x 40,000 dense LP (BASELINE config 3; objective with `--pos-permille` positive
       coefficients so the first-positive rule terminates in O(10^3) pivots) solved to optimality
       on the GPU and by the binary64 C twin with all host threads: full pivot-sequence comparison.
usage: python tools/parity_full_solves.py c2|c4 [--pos-permille P] [--out file.json]
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
from oracle import tier_d, tier_f  # noqa: E402


def digest(log):
    return hashlib.sha256(np.asarray(log, dtype=np.int32).tobytes()).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["c2", "c4"])
    ap.add_argument("--pos-permille", type=int, default=10)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    threads = tier_f.lib().tf_max_threads()
    rec = {"which": a.which, "seed": a.seed, "host_threads": threads}
    if a.which == "c2":
        m = n = 1000
        A, b, c = tier_f.gen_dense_feasible(m, n, a.seed)
        st = L.LPState(A, b, c, m, n)
        t0 = time.perf_counter()
        r = st.run()
        rec["gpu"] = {"verdict": int(r.verdict), "pivots": int(r.npivots), "v": st.v, "device_ms": r.device_ms,
                      "wall_s": time.perf_counter() - t0, "log_sha256": digest(st.pivot_log)}
        ref = tier_d.TierDState(A, b, c, nthreads=threads)
        t0 = time.perf_counter()
        status, k = ref.run()
        rA, rb, rc, rv, rpos = ref.read()
        glog = st.pivot_log
        first_diff = next((i for i, (x, y) in enumerate(zip(glog, ref.log)) if x != y), None)
        rec["reference_arithmetic"] = {"oracle": "oracle/tier_d.c (BigDecimal 15 digits HALF_UP restated)",
                                       "status": int(status), "pivots": int(k), "v": str(ref.v),
                                       "wall_s": time.perf_counter() - t0, "log_sha256": digest(ref.log)}
        rec["identical_sequence"] = glog == ref.log
        rec["first_divergence"] = first_diff
        rec["rel_err_v"] = abs(st.v - rv) / max(1.0, abs(rv))
        rec["max_rel_err_b"] = float(np.max(np.abs(st.b - rb) / np.maximum(1.0, np.abs(rb))))
    else:
        m, n = 20000, 40000
        st = L.LPState.synthetic_dense(m, n, a.seed, a.pos_permille)
        t0 = time.perf_counter()
        r = st.run()
        glog = st.pivot_log
        rec["pos_permille"] = a.pos_permille
        rec["gpu"] = {"verdict": int(r.verdict), "pivots": int(r.npivots), "v": st.v, "device_ms": r.device_ms,
                      "wall_s": time.perf_counter() - t0, "log_sha256": digest(glog)}
        rec["gpu"]["b_sha256"] = hashlib.sha256(st.b.tobytes()).hexdigest()
        rec["primal_nonzeros"] = int(np.count_nonzero(st.primal(n)))
        rec["note"] = ("the CPU twin of this solve takes ~20 min of host time, so it runs separately without "
                       "holding a GPU: tools/cpu_twin_c4.py prints the same digests")
    print(json.dumps(rec, indent=1))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(rec, f, indent=1)


if __name__ == "__main__":
    main()
