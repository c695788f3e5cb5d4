"""Time the tableau-update kernel variants on a synthetic dense LP (GPU box only).
usage: python tools/tune_update.py [m n pivots] [variants...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import linear_programming_solver_b200 as L


def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
    pivots = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    variants = [int(x) for x in sys.argv[4:]] or list(range(8))
    out = []
    for var in variants:
        # variant >= 100 selects the persistent loop kernel
        kw = dict(loop_mode=var - 98) if var >= 100 else dict(update_variant=var, loop_mode=1)
        st = L.LPState.synthetic_dense(m, n, 0, 1000, time_kernels=True, **kw)
        st.run(10)  # warm-up
        r = st.run(pivots)
        bytes_pp = st.algorithmic_bytes_per_pivot()
        upd_ms = r.update_ms / max(r.update_launches, 1)
        rec = dict(variant=var, m=m, n=n, pivots=int(r.npivots), ms_per_pivot=r.device_ms / max(r.npivots, 1),
                   update_ms=upd_ms, update_gbs=bytes_pp / upd_ms / 1e6 if upd_ms else None,
                   loop_gbs=bytes_pp * r.npivots / r.device_ms / 1e6, verdict=int(r.verdict))
        print(json.dumps(rec), flush=True)
        out.append(rec)
        st.close()
    return out


if __name__ == "__main__":
    main()
