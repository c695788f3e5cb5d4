"""One-line summary of bench.py JSON lines: python tools/bench_summary.py file.json [...]"""
import json
import sys

for path in sys.argv[1:]:
    try:
        lines = [l for l in open(path) if l.startswith("{")]
        d = json.loads(lines[-1])
    except Exception as ex:  # noqa: BLE001
        print(path, "unreadable:", ex)
        continue
    r = d.get("roofline", {})
    per = 1e6 / d["value"]
    ppl = r.get("pivots_per_launch", 1) or 1
    pass_us = r.get("avg_ms", 0) * 1e3 / ppl
    e2e = d.get("e2e", {}).get("value")
    # round-2 lines: `achieved` / `frac` are the physical DRAM rate of the step (round 1 had them in dram_*); in the
    # look-ahead loop a step holds the pass AND the next block's panel, so "outside the step" is launch overhead only
    dram, frac = r.get("dram_achieved", r.get("achieved", 0)), r.get("dram_frac", r.get("frac", 0))
    fp = (r.get("fp64") or {}).get("frac")
    par = (d.get("parity") or {}).get("ok")
    print("%s: N=%d  %.0f pivots/s  %.1f us/pivot = step %.1f + outside %.1f  (step %.3f ms, dram %.0f GB/s = %.2f of peak, fp64 %s)  "
          "e2e %s  clocks %s  parity %s"
          % (path.split("/")[-1], d["n_gpus"], d["value"], per, pass_us, per - pass_us, r.get("avg_ms", 0), dram, frac,
             "%.2f" % fp if fp else None, "%.0f" % e2e if e2e else None, d.get("clocks", {}).get("sm_mhz"), par))
