"""Parity of the CUDA path (through the C ABI) with the oracle.  Needs a B200: `-m gpu`.

Contract:
  P1  GPU == Tier F (binary64 twin) bit for bit: pivot sequence, verdict, every tableau cell.
  P2  GPU vs Tier D (the reference's decimal-15 arithmetic): identical (e,l) sequence on the
      non-degenerate continuous and the exact-integer families, same verdicts, objective and
      primal values within 1e-9 relative.
The first half mirrors the reference's own Spock specs (tests/golden/spock_vectors.py).
"""
import json
import os

import numpy as np
import pytest

from oracle import tier_f
from oracle.arith import Dec15, F64
from oracle.lp_text import LPInputReader as OracleReader
from oracle.simplex_ref import LPException as OLPException
from oracle.simplex_ref import LPSolver as OracleSolver
from oracle.simplex_ref import LPStandardForm as OracleForm
from oracle.simplex_ref import SolutionException as OSolutionException
from oracle.simplex_ref import primal_solution
from tests.golden import spock_vectors as G

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "input_txt_lps.json")) as _f:
    INPUT_LPS = json.load(_f)["lps"]

REL_TOL = 1e-9   # north star: objective and primal values within 1e-9 relative


def _pkg():
    import linear_programming_solver_b200 as L
    return L


def _f(x):
    return np.array([float(v) for v in x], dtype=np.float64)


def _f2(rows):
    return np.array([[float(x) for x in r] for r in rows], dtype=np.float64)


def _names(k):
    return ({i: "x%d" % (i + 1) for i in range(k)}, {"x%d" % (i + 1): i for i in range(k)})


# ---- LPStateSpec ----------------------------------------------------------------------------
@pytest.mark.parametrize("c,entering", G.GET_ENTERING)
def test_get_entering(c, entering):
    L = _pkg()
    st = L.LPState(np.zeros((0, len(c))), [], _f(c), 0, len(c))
    assert st.get_entering() == entering


@pytest.mark.parametrize("entering,leaving", G.GET_LEAVING["cases"])
def test_get_leaving(entering, leaving):
    L = _pkg()
    st = L.LPState(_f2(G.GET_LEAVING["A"]), _f(G.GET_LEAVING["b"]), np.zeros(4), 4, 4)
    assert st.get_leaving(entering) == leaving


def test_get_leaving_rejects_bad_index():
    L = _pkg()
    st = L.LPState(_f2(G.GET_LEAVING["A"]), _f(G.GET_LEAVING["b"]), np.zeros(4), 4, 4)
    with pytest.raises(ValueError):          # Validate.isTrue -> IllegalArgumentException
        st.get_leaving(4)
    with pytest.raises(ValueError):
        st.get_leaving(-1)


def _check_pivot(vec, case):
    L = _pkg()
    m, n = vec["m"], vec["n"]
    variables, coefficients = _names(m + n)
    st = L.LPState(_f2(vec["A"]), _f(vec["b"]), _f(vec["c"]), m, n, variables=variables,
                   coefficients=coefficients)
    st.pivot(case["e"], case["l"])
    assert np.array_equal(st.A, _f2(case["resA"]))
    assert np.array_equal(st.b, _f(case["resB"]))
    assert np.array_equal(st.c, _f(case["resC"]))
    assert st.v == float(case["resV"])
    assert st.variables == case["resVariables"]
    assert st.coefficients == case["resCoefficients"]


def test_pivot_1x1():
    _check_pivot(G.PIVOT_1x1, G.PIVOT_1x1)


@pytest.mark.parametrize("vec", [G.PIVOT_2x2, G.PIVOT_4x5, G.PIVOT_7x2], ids=["2x2", "4x5", "7x2"])
def test_pivot_vectors(vec):
    for case in vec["cases"]:
        _check_pivot(vec, case)


def test_zero_pivot_is_an_error():
    L = _pkg()
    st = L.LPState([[0.0, 1.0]], [1.0], [1.0, 1.0], 1, 2)
    with pytest.raises(ValueError):          # ArithmeticException in the reference (divide by zero)
        st.pivot(0, 0)


# ---- LPSolverSpec ---------------------------------------------------------------------------
@pytest.mark.parametrize("b,answer", G.MIN_IN_B)
def test_min_in_b(b, answer):
    assert _pkg().LPSolver.min_in_b(b) == answer


def test_aux_construction():
    L = _pkg()
    v = G.AUX_CONSTRUCTION
    variables, coefficients = _names(v["n"])
    form = L.LPStandardForm(v["A"], v["b"], v["c"], v["m"], v["n"], True, variables, coefficients)
    st = L.LPSolver().convert_into_aux_lp(form)
    assert np.array_equal(st.A, _f2(v["resA"]))
    assert np.array_equal(st.c, _f(v["resC"]))
    assert len(st.coefficients) == len(st.variables)
    assert "x0" in st.coefficients and "x0" in st.variables.values()


def _gpu_solve(case, fix=False):
    L = _pkg()
    form = L.LPStandardForm(case["A"], case["b"], case["c"], case["m"], case["n"], case["maximize"])
    solver = L.LPSolver(fix_restore_index=fix)
    verdict, value, message = "optimal", None, None
    try:
        value = solver.solve(form)
    except L.SolutionException as ex:
        verdict, message = "unbounded", str(ex)
    except L.LPException as ex:
        verdict, message = "infeasible", str(ex)
    except IndexError as ex:
        verdict, message = "index_error", str(ex)
    return solver, verdict, value, message


@pytest.mark.parametrize("case", G.SOLVE, ids=lambda c: c["name"])
def test_solve_known_answers(case):
    solver, verdict, value, message = _gpu_solve(case)
    assert verdict == case["verdict"]
    if verdict == "optimal":
        assert str(value) == case["value"]
    else:
        assert message == case["message"]
    # index-level pivot sequence against the Tier-D oracle run on the same LP
    oform = OracleForm(case["A"], case["b"], case["c"], case["m"], case["n"], case["maximize"], arith=Dec15)
    osolver = OracleSolver(Dec15)
    try:
        osolver.solve(oform)
    except OLPException:
        pass
    assert solver.info.phase1_log == osolver.trace.phase1_log
    if verdict != "infeasible":
        assert solver.info.phase2_log == osolver.trace.phase2_log
    if case.get("x0_index") is not None:
        assert solver.info.x0_index == case["x0_index"]


def test_exception_hierarchy():
    L = _pkg()
    assert issubclass(L.SolutionException, L.LPException)   # SolutionException.java:3
    case = [c for c in G.SOLVE if c["name"] == "unbounded_after_phase1"][0]
    form = L.LPStandardForm(case["A"], case["b"], case["c"], case["m"], case["n"], True)
    with pytest.raises(L.LPException) as ei:                 # LPSolverSpec.groovy:175 catches LPException
        L.LPSolver().solve(form)
    assert str(ei.value) == "This linear program is unbounded"


def test_min_negates_c_in_place():
    L = _pkg()
    case = [c for c in G.SOLVE if c["name"] == "minimization"][0]
    form = L.LPStandardForm(case["A"], case["b"], case["c"], case["m"], case["n"], False)
    L.LPSolver().solve(form)
    assert form.c.tolist() == [3.0, -1.0]                    # LPSolver.java:86-89


def test_aux_lp_solving():
    L = _pkg()
    v = G.AUX_SOLVE
    st = L.LPState(_f2(v["A"]), _f(v["b"]), _f(v["c"]), v["m"], v["n"])
    st.pivot(v["index_of_x0"], v["min_in_b"])                # LPSolver.java:138
    res = st.run()
    assert res.verdict == 1
    assert st.v == float(v["resV"])
    assert st.position_of(v["index_of_x0"]) == v["x0_index"]


def test_restore_initial_lp():
    L = _pkg()
    v = G.RESTORE
    st = L.LPState(_f2(v["A"]), _f(v["b"]), _f(v["c"]), v["m"], v["n"])
    # x1 is basic in row 3, x2 in row 1 (positions 6 and 4 of the aux map, n_aux = 3)
    st.drop_column(v["index_of_x0"])
    st.rebuild_objective([(0, 3, 1.0), (0, 1, 1.0)])
    assert np.array_equal(st.A, _f2(v["resA"]))
    assert np.array_equal(st.b, _f(v["resB"]))
    assert np.array_equal(st.c, _f(v["resC"]))
    assert st.v == float(v["resV"])
    assert st.n == 2 and st.m == 4


# ---- io_files/input.txt ---------------------------------------------------------------------
@pytest.mark.parametrize("entry", INPUT_LPS, ids=lambda e: "lp%d" % e["index"])
@pytest.mark.parametrize("fix", [False, True], ids=["asref", "fixed"])
def test_input_txt(entry, fix):
    L = _pkg()
    want_f = entry["f64_%s" % ("fixed" if fix else "asref")]
    want_d = entry["dec15_%s" % ("fixed" if fix else "asref")]
    if want_f["verdict"] == "parse_error":
        pytest.skip("unparsable LP (reference regex rejects '+-x4')")
    form = OracleReader(F64).read_lp(entry["text"])
    case = dict(A=form.A, b=form.b, c=form.c, m=form.m, n=form.n, maximize=form.maximize)
    solver, verdict, value, message = _gpu_solve(case, fix=fix)
    # P1: bit-level twin
    assert verdict == want_f["verdict"]
    assert [list(p) for p in solver.info.phase1_log] == want_f["phase1_log"]
    if verdict != "infeasible":
        assert [list(p) for p in solver.info.phase2_log] == want_f["phase2_log"]
    if verdict == "optimal":
        assert str(value) == want_f["value"]
        assert np.allclose(solver.info.primal, [float(x) for x in want_f["primal"]], rtol=REL_TOL, atol=1e-12)
    # P2: reference arithmetic — verdict and 6-decimal objective whenever both took the same path
    if want_d["phase1_log"] == want_f["phase1_log"] and want_d["phase2_log"] == want_f["phase2_log"]:
        assert verdict == want_d["verdict"]
        if verdict == "optimal":
            assert str(value) == want_d["value"]


def test_input_txt_lp1_fixture():
    """BASELINE config 0: io_files/input.txt LP #1 -> 7.000 with the primal of output.txt:214-233."""
    L = _pkg()
    form = OracleReader(F64).read_lp_file_text("\n\n".join(e["text"] for e in INPUT_LPS))
    f = L.LPStandardForm(form.A, form.b, form.c, form.m, form.n, form.maximize)
    solver = L.LPSolver()
    assert str(solver.solve(f)) == G.INPUT_TXT_LP1["value"]
    assert solver.info.primal.tolist() == [float(x) for x in G.INPUT_TXT_LP1["primal"]]
    assert len(solver.info.phase2_log) == 16


# ---- synthetic families -----------------------------------------------------------------------
def test_device_generator_matches_oracle():
    L = _pkg()
    for (m, n, seed, pp) in [(7, 5, 0, 1000), (33, 70, 1, 1000), (64, 100, 2, 100)]:
        st = L.LPState.synthetic_dense(m, n, seed, pp)
        A, b, c = tier_f.gen_dense_feasible(m, n, seed, pp)
        assert np.array_equal(st.A, A) and np.array_equal(st.b, b) and np.array_equal(st.c, c)
        assert st.v == 0.0


@pytest.mark.parametrize("m,n,seed", [(5, 7, 0), (12, 9, 1), (20, 20, 2), (40, 80, 3), (100, 60, 4),
                                      (150, 150, 5), (257, 1030, 6), (300, 300, 7)])
def test_dense_feasible_bit_exact_vs_tier_f(m, n, seed):
    """P1 on the C2 family: full solve, every cell of the final tableau identical."""
    L = _pkg()
    A, b, c = tier_f.gen_dense_feasible(m, n, seed)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy())
    status, k = ref.run()
    st = L.LPState(A, b, c, m, n)
    res = st.run()
    assert res.verdict == {tier_f.OPTIMAL: 1, tier_f.UNBOUNDED: 2}[status]
    assert res.npivots == k
    assert st.pivot_log == ref.log
    assert np.array_equal(st.A, ref.A)
    assert np.array_equal(st.b, ref.b)
    assert np.array_equal(st.c, ref.c)
    assert st.v == ref.v[0]
    assert np.array_equal(st.positions, ref.pos2var)


@pytest.mark.parametrize("m,n,seed", [(20, 20, 0), (30, 60, 1), (60, 40, 2)])
def test_dense_feasible_vs_tier_d(m, n, seed):
    """P2 on the non-degenerate continuous family: same (e,l) sequence as the reference's
    decimal-15 arithmetic, objective and primal within 1e-9 relative."""
    L = _pkg()
    A, b, c = tier_f.gen_dense_feasible(m, n, seed)
    oform = OracleForm(A.tolist(), b.tolist(), c.tolist(), m, n, True, arith=Dec15)
    osolver = OracleSolver(Dec15)
    oval = osolver.solve(oform)
    form = L.LPStandardForm(A, b, c, m, n, True)
    solver = L.LPSolver()
    val = solver.solve(form)
    assert solver.info.phase2_log == osolver.trace.phase2_log
    ref_v = float(osolver.trace.raw_v)
    assert abs(solver.info.raw_value - ref_v) <= REL_TOL * max(1.0, abs(ref_v))
    ox = [float(x) for x in primal_solution_from_positions(osolver.trace.final_state, n)]
    assert np.allclose(solver.info.primal, ox, rtol=REL_TOL, atol=REL_TOL)
    assert str(val) == str(oval)


def primal_solution_from_positions(state, n):
    # the no-names oracle state has no maps; rebuild positions from its pivot log
    pos = list(range(state.n + state.m))
    for e, l in state.pivot_log:
        pos[e], pos[state.n + l] = pos[state.n + l], pos[e]
    x = [0] * n
    for p in range(state.n, state.n + state.m):
        if pos[p] < n:
            x[pos[p]] = state.b[p - state.n]
    return x


def test_step_by_step_equals_run():
    """Driving getEntering/getLeaving/pivot from the host gives the same state as the on-device loop."""
    L = _pkg()
    A, b, c = tier_f.gen_dense_feasible(30, 40, 11)
    s1 = L.LPState(A, b, c, 30, 40)
    s2 = L.LPState(A, b, c, 30, 40)
    while True:
        e = s1.get_entering()
        if e == -1:
            break
        l = s1.get_leaving(e)
        assert l != -1
        s1.pivot(e, l)
    r = s2.run()
    assert r.verdict == 1
    assert s1.pivot_log == s2.pivot_log
    assert np.array_equal(s1.A, s2.A) and np.array_equal(s1.b, s2.b) and np.array_equal(s1.c, s2.c)
    assert s1.v == s2.v


def test_pivot_cap_and_resume():
    L = _pkg()
    A, b, c = tier_f.gen_dense_feasible(60, 60, 3)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy())
    ref.run()
    st = L.LPState(A, b, c, 60, 60)
    r = st.run(10)
    assert r.verdict == 3 and r.npivots == 10
    assert st.pivot_log == ref.log[:10]
    r = st.run(0)
    assert r.verdict == 3 and r.npivots == 0
    r = st.run()
    assert r.verdict == 1 and r.total_pivots == len(ref.log)
    assert st.pivot_log == ref.log
    assert np.array_equal(st.A, ref.A)


def _random_phase1(seed):
    rng = np.random.default_rng(seed)
    m, n = int(rng.integers(4, 40)), int(rng.integers(3, 30))
    A = rng.integers(-4, 9, size=(m, n)).astype(np.float64)
    xs = rng.integers(0, 4, size=n).astype(np.float64)
    b = A @ xs + rng.integers(0, 3, size=m)
    flip = rng.random(m) < 0.4
    A[flip] *= -1
    b[flip] *= -1
    if not (b < 0).any():
        b[0] = -abs(b[0]) - 1
    c = rng.integers(-3, 6, size=n).astype(np.float64)
    return A, b, c, m, n


@pytest.mark.parametrize("seed", range(24))
@pytest.mark.parametrize("fix", [False, True], ids=["asref", "fixed"])
def test_random_phase1_vs_tier_f(seed, fix):
    """P1 through phase 1: aux LP, forced pivot, degenerate pivot, column drop, objective
    rebuild, phase 2 — sequences, verdict and value identical to the binary64 twin."""
    A, b, c, m, n = _random_phase1(seed)
    r = tier_f.solve(A, b, c, True, fix_restore_index=fix)
    solver, verdict, value, message = _gpu_solve(dict(A=A, b=b, c=c, m=m, n=n, maximize=True), fix=fix)
    assert verdict == r.verdict
    assert solver.info.phase1_log == r.phase1_log
    if verdict in ("optimal", "unbounded"):
        assert solver.info.phase2_log == r.phase2_log
    if verdict == "optimal":
        assert solver.info.raw_value == r.value
        assert np.array_equal(solver.info.primal, r.primal)


def test_mid_size_capped_bit_exact():
    """1,000 x 1,000 (BASELINE config 1), first 400 pivots: sequence and every cell identical."""
    L = _pkg()
    m = n = 1000
    A, b, c = tier_f.gen_dense_feasible(m, n, 0)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy(), nthreads=tier_f.lib().tf_max_threads())
    status, k = ref.run(400)
    st = L.LPState(A, b, c, m, n)
    r = st.run(400)
    assert r.npivots == k == 400 and r.verdict == 3 and status == tier_f.PIVOT_CAP
    assert st.pivot_log == ref.log
    assert np.array_equal(st.A, ref.A) and np.array_equal(st.b, ref.b) and np.array_equal(st.c, ref.c)
    assert st.v == ref.v[0]


def test_wide_tableau_capped_bit_exact():
    """A tableau wider than one CTA column chunk with a ragged edge (n+1 not a multiple of 4)."""
    L = _pkg()
    m, n = 700, 5001
    A, b, c = tier_f.gen_dense_feasible(m, n, 2)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy(), nthreads=tier_f.lib().tf_max_threads())
    ref.run(60)
    st = L.LPState(A, b, c, m, n)
    st.run(60)
    assert st.pivot_log == ref.log
    assert np.array_equal(st.A, ref.A) and np.array_equal(st.b, ref.b) and np.array_equal(st.c, ref.c)


def test_unbounded_column_and_immediate_verdicts():
    L = _pkg()
    m, n = 50, 40
    A, b, c = tier_f.gen_dense_feasible(m, n, 5)
    A[:, 0] = -A[:, 0]                      # column 0: c>0 and A<=0 -> unbounded at once
    st = L.LPState(A, b, c, m, n)
    r = st.run()
    assert r.verdict == 2 and r.npivots == 0 and r.last_entering == 0
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy())
    assert ref.run()[0] == tier_f.UNBOUNDED
    c2 = -np.abs(c)                         # nothing to enter -> optimal at once
    st = L.LPState(A, b, c2, m, n)
    r = st.run()
    assert r.verdict == 1 and r.npivots == 0 and st.v == 0.0


def test_degenerate_ties_lowest_row_wins():
    """Exact-integer degenerate family (bipartite incidence, b = 1): many zero-ratio ties;
    sequence identical to Tier D and Tier F."""
    L = _pkg()
    rng = np.random.default_rng(0)
    left, right, n = 12, 10, 60
    A = np.zeros((left + right, n))
    for j in range(n):
        A[rng.integers(0, left), j] = 1.0
        A[left + rng.integers(0, right), j] = 1.0
    b = np.ones(left + right)
    c = np.ones(n)
    m = left + right
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy())
    ref.run()
    oform = OracleForm(A.tolist(), b.tolist(), c.tolist(), m, n, True, arith=Dec15)
    osolver = OracleSolver(Dec15)
    oval = osolver.solve(oform)
    st = L.LPState(A, b, c, m, n)
    r = st.run()
    assert r.verdict == 1
    assert st.pivot_log == ref.log == osolver.trace.phase2_log
    assert st.v == float(osolver.trace.raw_v)
    assert np.array_equal(st.A, ref.A)


# ---- parity with the reference's own arithmetic at BASELINE sizes (C decimal-15 oracle) ----------
@pytest.mark.parametrize("m,n,seed,cap", [(300, 300, 0, -1), (250, 500, 1, -1), (1000, 1000, 0, 1200)])
def test_sequence_vs_reference_arithmetic_mid_size(m, n, seed, cap):
    """P2 at scale: the GPU's (entering, leaving) sequence equals the one BigDecimal(15, HALF_UP)
    arithmetic produces (oracle/tier_d.c), objective and every b within 1e-9 relative.
    1000 x 1000 is BASELINE config 1 (first 1200 pivots)."""
    from oracle import tier_d
    L = _pkg()
    A, b, c = tier_f.gen_dense_feasible(m, n, seed)
    ref = tier_d.TierDState(A, b, c, nthreads=tier_f.lib().tf_max_threads())
    status, k = ref.run(cap)
    st = L.LPState(A, b, c, m, n)
    r = st.run(cap)
    assert r.npivots == k
    assert r.verdict == {tier_d.OPTIMAL: 1, tier_d.UNBOUNDED: 2, tier_d.PIVOT_CAP: 3}[status]
    assert st.pivot_log == ref.log
    rA, rb, rc, rv, rpos = ref.read()
    assert abs(st.v - rv) <= REL_TOL * max(1.0, abs(rv))
    assert np.allclose(st.b, rb, rtol=REL_TOL, atol=REL_TOL)
    assert np.array_equal(st.positions, rpos)


# ---- C3 family: mixed <= / >= / == rows, phase 1 forced --------------------------------------
@pytest.mark.parametrize("m,n,weq", [(60, 60, True), (100, 100, False), (300, 300, True), (300, 300, False),
                                      (200, 400, True)])
@pytest.mark.parametrize("fix", [False, True], ids=["asref", "fixed"])
def test_mixed_rows_phase1_vs_tier_f(m, n, weq, fix):
    """BASELINE config 2 at reduced size (the full 10,000 x 10,000 instance needs millions of pivots
    under the first-positive rule): aux LP on the device, forced pivot, loop, degenerate pivot,
    column drop, objective rebuild, phase 2 — P1: every index and the value equal the twin's."""
    A, b, c = tier_f.gen_mixed_rows(m, n, 0, weq)
    assert (b < 0).any()
    r = tier_f.solve(A, b, c, True, fix_restore_index=fix, nthreads=4)
    solver, verdict, value, message = _gpu_solve(dict(A=A, b=b, c=c, m=m, n=n, maximize=True), fix=fix)
    assert verdict == r.verdict
    assert solver.info.used_phase1
    assert solver.info.x0_index == r.x0_index
    assert solver.info.phase1_log == r.phase1_log
    assert solver.info.phase2_log == r.phase2_log
    if verdict == "optimal":
        assert solver.info.raw_value == r.value
        assert np.array_equal(solver.info.primal, r.primal)
        if fix:      # with the index shift the answer is a true optimum: primal feasible, objective = c.x
            x = solver.info.primal
            assert (x >= -1e-9).all() and (A @ x <= b + 1e-6).all()
            assert abs(c @ x - solver.info.raw_value) <= 1e-7 * max(1.0, abs(solver.info.raw_value))


@pytest.mark.parametrize("m,n", [(30, 20), (60, 60)])
def test_mixed_rows_without_equalities_vs_tier_d(m, n):
    """P2 through phase 1 on the variant without '==' rows (SURVEY §7 hard part 1): same (e,l)
    sequence as the reference's decimal arithmetic in both phases, objective within 1e-9."""
    L = _pkg()
    A, b, c = tier_f.gen_mixed_rows(m, n, 1, with_equalities=False)
    oform = OracleForm(A.tolist(), b.tolist(), c.tolist(), m, n, True, arith=Dec15)
    oform.key_order = "index"
    osolver = OracleSolver(Dec15, fix_restore_index=True)
    oval = osolver.solve(oform)
    solver = L.LPSolver(fix_restore_index=True)
    val = solver.solve(L.LPStandardForm(A, b, c, m, n, True))
    assert solver.info.phase1_log == osolver.trace.phase1_log
    assert solver.info.phase2_log == osolver.trace.phase2_log
    ref_v = float(osolver.trace.raw_v)
    assert abs(solver.info.raw_value - ref_v) <= REL_TOL * max(1.0, abs(ref_v))
    assert str(val) == str(oval)
