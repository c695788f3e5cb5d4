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
    print("%s: N=%d  %.0f pivots/s  %.1f us/pivot = pass %.1f + panel %.1f  (pass %.3f ms, dram %.0f GB/s = %.2f of peak)  e2e %s  clocks %s"
          % (path.split("/")[-1], d["n_gpus"], d["value"], per, pass_us, per - pass_us, r.get("avg_ms", 0),
             r.get("dram_achieved", 0), r.get("dram_frac", 0), "%.0f" % e2e if e2e else None, d.get("clocks", {}).get("sm_mhz")))
