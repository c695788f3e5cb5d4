"""The C Tier-F oracle (oracle/tier_f.c) against the Python restatement (oracle/simplex_ref.py
with binary64 arithmetic) and the golden vectors.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import tier_f
from oracle.arith import F64
from oracle.lp_text import LPInputReader
from oracle.simplex_ref import LPException, LPSolver, LPStandardForm, LPState, SolutionException
from tests.golden import spock_vectors as G

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "input_txt_lps.json")) as _f:
    INPUT_LPS = json.load(_f)["lps"]


def _f(x):
    return np.array(x, dtype=np.float64)


def _f2(rows):
    return np.array([[float(x) for x in r] for r in rows], dtype=np.float64)


@pytest.mark.parametrize("c,entering", G.GET_ENTERING)
def test_get_entering(c, entering):
    st = tier_f.TierFState(np.zeros((0, len(c))), [], _f(c))
    assert st.get_entering() == entering


@pytest.mark.parametrize("entering,leaving", G.GET_LEAVING["cases"])
def test_get_leaving(entering, leaving):
    st = tier_f.TierFState(_f2(G.GET_LEAVING["A"]), _f(G.GET_LEAVING["b"]), np.zeros(4))
    assert st.get_leaving(entering) == leaving


@pytest.mark.parametrize("nthreads", [1, 4])
@pytest.mark.parametrize("vec", [G.PIVOT_2x2, G.PIVOT_4x5, G.PIVOT_7x2], ids=["2x2", "4x5", "7x2"])
def test_pivot_vectors(vec, nthreads):
    for case in vec["cases"]:
        st = tier_f.TierFState(_f2(vec["A"]), _f(vec["b"]), _f(vec["c"]), nthreads=nthreads)
        st.pivot(case["e"], case["l"])
        assert np.array_equal(st.A, _f2(case["resA"]))
        assert np.array_equal(st.b, _f([float(x) for x in case["resB"]]))
        assert np.array_equal(st.c, _f([float(x) for x in case["resC"]]))
        assert st.v[0] == float(case["resV"])
        # positions: names x1.. map to ids 0..
        want = [int(case["resVariables"][p][1:]) - 1 for p in range(vec["m"] + vec["n"])]
        assert st.pos2var.tolist() == want


@pytest.mark.parametrize("b,answer", G.MIN_IN_B)
def test_min_in_b(b, answer):
    assert tier_f.min_in_b(_f(b)) == answer


@pytest.mark.parametrize("case", G.SOLVE, ids=lambda c: c["name"])
def test_solve_known_answers(case):
    r = tier_f.solve(case["A"], case["b"], case["c"], case["maximize"])
    assert r.verdict == case["verdict"]
    if r.verdict == "optimal":
        assert str(F64.set_scale6(r.value)) == case["value"]
    else:
        assert r.message == case["message"]
    if case.get("x0_index") is not None:
        assert r.x0_index == case["x0_index"]


@pytest.mark.parametrize("entry", INPUT_LPS, ids=lambda e: "lp%d" % e["index"])
@pytest.mark.parametrize("fix", [False, True], ids=["asref", "fixed"])
def test_input_txt(entry, fix):
    want = entry["f64_%s" % ("fixed" if fix else "asref")]
    if want["verdict"] == "parse_error":
        pytest.skip("unparsable LP")
    form = LPInputReader(F64).read_lp(entry["text"])
    r = tier_f.solve(form.A, form.b, form.c, form.maximize, fix_restore_index=fix)
    assert r.verdict == want["verdict"]
    assert [list(p) for p in r.phase1_log] == want["phase1_log"]
    assert [list(p) for p in r.phase2_log] == want["phase2_log"]
    if r.verdict == "optimal":
        assert str(F64.set_scale6(r.value)) == want["value"]
        assert np.allclose(r.primal, [float(x) for x in want["primal"]], rtol=0, atol=1e-12)


def _py_f64_solve(A, b, c, fix=False, key_order="index"):
    m, n = A.shape
    form = LPStandardForm(A.tolist(), b.tolist(), c.tolist(), m, n, True, arith=F64)
    form.key_order = key_order
    s = LPSolver(F64, fix_restore_index=fix)
    try:
        val = s.solve(form)
        return "optimal", s
    except SolutionException:
        return "unbounded", s
    except LPException:
        return "infeasible", s
    except IndexError:
        return "index_error", s


@pytest.mark.parametrize("m,n,seed", [(5, 7, 0), (12, 9, 1), (20, 20, 2), (30, 45, 3), (40, 25, 4)])
@pytest.mark.parametrize("nthreads", [1, 3])
def test_random_feasible_bit_exact_vs_python(m, n, seed, nthreads):
    A, b, c = tier_f.gen_dense_feasible(m, n, seed)
    verdict, s = _py_f64_solve(A, b, c)
    r = tier_f.solve(A, b, c, True, nthreads=nthreads)
    assert r.verdict == verdict == "optimal"
    assert r.phase2_log == s.trace.phase2_log
    fs = s.trace.final_state
    assert np.array_equal(r.state.A, np.array(fs.A))          # every cell bit-identical
    assert np.array_equal(r.state.b, np.array(fs.b))
    assert np.array_equal(r.state.c, np.array(fs.c))
    assert r.state.v[0] == fs.v


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("fix", [False, True])
def test_random_phase1_bit_exact_vs_python(seed, fix):
    rng = np.random.default_rng(seed)
    m, n = int(rng.integers(4, 14)), int(rng.integers(3, 12))
    A = rng.integers(-4, 9, size=(m, n)).astype(np.float64)
    xs = rng.integers(0, 4, size=n).astype(np.float64)
    b = A @ xs + rng.integers(0, 3, size=m)
    flip = rng.random(m) < 0.4                      # '>=' rows lowered by negation => some b < 0
    A[flip] *= -1
    b[flip] *= -1
    if not (b < 0).any():
        b[0] = -abs(b[0]) - 1
    c = rng.integers(-3, 6, size=n).astype(np.float64)
    verdict, s = _py_f64_solve(A, b, c, fix=fix)
    r = tier_f.solve(A, b, c, True, fix_restore_index=fix)
    assert r.verdict == verdict
    assert r.phase1_log == s.trace.phase1_log
    if verdict in ("optimal", "unbounded"):
        assert r.phase2_log == s.trace.phase2_log
    if verdict == "optimal":
        assert r.value == s.trace.raw_v


def test_generator_matches_c():
    L = tier_f.lib()
    ks = np.array([0, 1, 2, 12345, 2 ** 40 + 7], dtype=np.uint64)
    for seed in (0, 1, 2):
        want = [L.tf_u(seed, int(k)) for k in ks]
        assert tier_f.u(seed, ks).tolist() == want
        assert all(0 < w <= 1 and (w * 2 ** 20) == int(w * 2 ** 20) for w in want)
