"""Time the blocked loop (pivots per tableau pass x pass-kernel tile shape) on a synthetic dense LP.
GPU box only.  usage: python tools/tune_blocked.py [m n blocks_per_run] [--blocks 8,16,32] [--variants 0,1,2]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import linear_programming_solver_b200 as L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("m", type=int, nargs="?", default=20000)
    ap.add_argument("n", type=int, nargs="?", default=40000)
    ap.add_argument("passes", type=int, nargs="?", default=6)
    ap.add_argument("--blocks", default="4,8,16,24,32")
    ap.add_argument("--variants", default="0")
    ap.add_argument("--mode", type=int, default=6, help="5 = two launches per pivot, 6 = one cooperative panel launch "
                    "per block + the pass, 7 = look-ahead loop (panel and pass side by side), 8 = look-ahead, both roles in every CTA")
    ap.add_argument("--panel", default="0", help="look-ahead loop: CTAs of the panel role (comma list, 0 = auto)")
    ap.add_argument("--chunk", default="0", help="TMA pass: rows per chunk (comma list, 0 = auto)")
    a = ap.parse_args()
    combos = [(S, var, P, ch) for S in [int(x) for x in a.blocks.split(",")] for var in [int(x) for x in a.variants.split(",")]
              for P in [int(x) for x in a.panel.split(",")] for ch in [int(x) for x in a.chunk.split(",")]]
    for S, var, P, ch in combos:
        if True:
            st = L.LPState.synthetic_dense(a.m, a.n, 0, 1000, time_kernels=True, loop_mode=a.mode, block_pivots=S,
                                           update_variant=var, panel_ctas=P, pass_chunk_rows=ch)
            st.run(2 * S)  # warm-up
            r = st.run(a.passes * S)
            bytes_pp = st.algorithmic_bytes_per_pivot()
            pass_ms = r.update_ms / max(r.update_launches, 1)
            rec = dict(block=S, variant=var, mode=a.mode, panel_ctas=P, chunk_rows=ch, m=a.m, n=a.n, pivots=int(r.npivots),
                       pivots_per_s=1e3 * r.npivots / r.device_ms, us_per_pivot=1e3 * r.device_ms / max(r.npivots, 1),
                       pass_ms=pass_ms, pass_dram_gbs=bytes_pp / pass_ms / 1e6 if pass_ms else None,
                       panel_us_per_pivot=1e3 * (r.device_ms - r.update_ms) / max(r.npivots, 1),
                       passes=int(r.update_launches), launches=int(r.kernel_launches), verdict=int(r.verdict))
            print(json.dumps(rec), flush=True)
            st.close()


if __name__ == "__main__":
    main()
