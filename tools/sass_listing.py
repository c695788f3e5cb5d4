"""Regenerate profiles/rNN_sass_*.txt and profiles/rNN_sass_summary.md from the built library
(cuobjdump -sass works without a GPU).  usage: python tools/sass_listing.py [round tag, default r02]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "linear_programming_solver_b200", "liblps_b200.so")

# (file tag, title, regex on the demangled name; the first match is listed)
KERNELS = [
    ("kb_sweep_shapeD", "kb_sweep<shape D> (the pass as a TMA + mbarrier pipeline, stand-alone)",
     r"lps::kb_sweep<lps::SweepShape<16, 4, 2, 8, 4, 2, false>"),
    ("kb_step_flush", "kb_step_flush<sharded, 4 lanes, 4 rows / group, 8 groups> (look-ahead step: panel role + cp.async pass role)",
     r"lps::kb_step_flush<true, 4, 4, 8, true>"),
    ("kb_step_shapeD", "kb_step<sharded, shape D> (look-ahead step: panel role + TMA pass role)",
     r"lps::kb_step<true, lps::SweepShape<16, 4, 2, 8, 4, 2, false>"),
    ("kb_flush", "kb_flush<4, 4, 16, prefetch> (the cp.async pass, stand-alone)", r"lps::kb_flush<4, 4, 16, true>"),
    ("kb_panel", "kb_panel<sharded> (serial blocked loop: the panel as a kernel of its own)", r"lps::kb_panel<true>"),
]
COUNTS = [("DMUL", r"\bDMUL\b"), ("DADD", r"\bDADD\b"), ("DFMA (inside __ddiv_rn only)", r"\bDFMA\b"), ("LDG", r"\bLDG\."),
          ("STG", r"\bSTG\."), ("LDS", r"\bLDS\b|\bLDS\."), ("STS", r"\bSTS\b|\bSTS\."), ("LDGSTS (cp.async)", r"\bLDGSTS"),
          ("UTMALDG (TMA load)", r"\bUTMALDG"), ("UBLKCP (bulk copy)", r"\bUBLKCP"), ("SYNCS (mbarrier)", r"\bSYNCS\."),
          ("NANOSLEEP.SYNCS", r"NANOSLEEP\.SYNCS"), ("ATOM/RED", r"\bATOMG|\bRED\.|\bATOMS"), ("BAR", r"\bBAR\."),
          ("CCTL (L2 prefetch)", r"\bCCTL"), ("LDL", r"\bLDL"), ("STL", r"\bSTL")]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    names = subprocess.run(["cuobjdump", "-elf", LIB], capture_output=True, text=True).stdout
    mangled = sorted(set(re.findall(r"\.text\.(_ZN[0-9A-Za-z_]+)", names)))
    demangled = subprocess.run(["c++filt"] + mangled, capture_output=True, text=True).stdout.splitlines()
    lines_md = ["# SASS listings of the round-2 kernels (cuobjdump -sass, sm_100a; made by tools/sass_listing.py)", ""]
    for ftag, title, rx in KERNELS:
        pick = [m for m, d in zip(mangled, demangled) if re.search(rx, d)]
        if not pick:
            lines_md.append("* `%s`: not in the library" % ftag)
            continue
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", pick[0], LIB], capture_output=True, text=True).stdout
        ins = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip() for l in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
        body = "\n".join(ins)
        counts = ", ".join("%s %d" % (n, len(re.findall(r, body))) for n, r in COUNTS)
        out = os.path.join(ROOT, "profiles", "%s_sass_%s.txt" % (tag, ftag))
        with open(out, "w") as f:
            f.write("# %s\n# cuobjdump -sass liblps_b200.so (sm_100a), function %s\n# %d instructions: %s\n" % (title, pick[0][:100], len(ins), counts))
            f.write(body + "\n")
        lines_md += ["* `profiles/%s_sass_%s.txt` — %s" % (tag, ftag, title), "  %d instructions: %s" % (len(ins), counts)]
    lines_md += ["", "No FMA touches a tableau value: DFMA appears only inside the IEEE division sequence `__ddiv_rn` (Newton steps on "
                 "the reciprocal); every update is a DMUL followed by a DADD, each rounded on its own (`-fmad=false`).  The TMA pass "
                 "shows `UTMALDG.2D` (cp.async.bulk.tensor), `SYNCS.*` (mbarrier arrive / expect_tx / try_wait) and `NANOSLEEP.SYNCS` "
                 "(the suspend-time hint of try_wait); the cp.async pass shows `LDGSTS`, `LDG.E.ENL2.256` / `STG.E.256` and "
                 "`CCTL.E.PF2` (prefetch.global.L2).", ""]
    with open(os.path.join(ROOT, "profiles", "%s_sass_summary.md" % tag), "w") as f:
        f.write("\n".join(lines_md))
    print("\n".join(lines_md))


if __name__ == "__main__":
    main()
