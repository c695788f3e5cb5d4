"""Generate tests/golden/input_txt_lps.json from the reference's own fixture file.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
Reads `io_files/input.txt` (15 blank-line-separated example LPs), parses each with the oracle's
restatement of LPInputReader, solves it with the Tier-D oracle (exact decimal-15 arithmetic)
both as the reference is written and with the restoreInitialLP index shift applied, and
records verdict, 6-decimal objective, pivot logs and primal values.  The LP texts travel in
the JSON so that nothing at test time needs /root/reference.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.arith import Dec15, F64  # noqa: E402
from oracle.lp_text import LPInputReader, split_lp_file  # noqa: E402
from oracle.simplex_ref import LPException, LPSolver, SolutionException, primal_solution  # noqa: E402

SRC = "/root/reference/io_files/input.txt"


def run(text, arith, fix):
    try:
        st = LPInputReader(arith).read_lp(text)
    except LPException as ex:
        return {"verdict": "parse_error", "message": str(ex)}
    names = [st.variables[i] for i in range(st.n)]
    n0 = st.n
    s = LPSolver(arith, fix_restore_index=fix)
    out = {"m": st.m, "n": st.n, "maximize": st.maximize}
    try:
        val = s.solve(st)
        out.update(verdict="optimal", value=str(val), raw_v=str(s.trace.raw_v))
        fs = s.trace.final_state
        x = primal_solution(fs, n0, name_of=lambda k: names[k])
        out["primal"] = [str(v) for v in x]
    except SolutionException as ex:
        out.update(verdict="unbounded" if "unbounded" in str(ex) else "error", message=str(ex))
    except LPException as ex:
        out.update(verdict="infeasible" if "infeasible" in str(ex) else "error", message=str(ex))
    except IndexError as ex:
        out.update(verdict="index_error", message=str(ex))
    out["phase1_log"] = [list(p) for p in s.trace.phase1_log]
    out["phase2_log"] = [list(p) for p in s.trace.phase2_log]
    out["x0_index"] = s.trace.x0_final_index
    return out


def main():
    text = open(SRC).read()
    lps = []
    for k, block in enumerate(split_lp_file(text)):
        entry = {"index": k + 1, "text": block}
        for arith in (Dec15, F64):
            for fix in (False, True):
                entry["%s_%s" % (arith.name, "fixed" if fix else "asref")] = run(block, arith, fix)
        lps.append(entry)
    with open(os.path.join(HERE, "input_txt_lps.json"), "w") as f:
        json.dump({"source": "io_files/input.txt", "lps": lps}, f, indent=1)
    print("wrote", len(lps), "LPs")


if __name__ == "__main__":
    main()
