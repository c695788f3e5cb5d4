"""Pin the oracle (both number systems) against every known-answer vector the reference's own
tests hold for the pivot path (tests/golden/spock_vectors.py), the recorded pivot traces in
logs/lp_solver.log and the io_files/input.txt fixture.  CPU only."""
import copy
import json
import os

import pytest

from oracle.arith import Dec15, F64
from oracle.lp_text import LPInputReader
from oracle.simplex_ref import (LPException, LPSolver, LPStandardForm, LPState, SolutionException,
                                primal_solution)
from tests.golden import spock_vectors as G

ARITHS = [Dec15, F64]
HERE = os.path.dirname(os.path.abspath(__file__))


def _conv(ar, xs):
    return [ar.conv(x) for x in xs]


def _conv2(ar, rows):
    return [[ar.conv(x) for x in r] for r in rows]


def _names(k):
    return ({i: "x%d" % (i + 1) for i in range(k)}, {"x%d" % (i + 1): i for i in range(k)})


def _eq(ar, got, want):
    return all(ar.cmp(g, ar.conv(w)) == 0 for g, w in zip(got, want)) and len(got) == len(want)


@pytest.mark.parametrize("ar", ARITHS, ids=lambda a: a.name)
@pytest.mark.parametrize("c,entering", G.GET_ENTERING)
def test_get_entering(ar, c, entering):
    st = LPState([[]], [], _conv(ar, c), 0, len(c), arith=ar)
    assert st.get_entering() == entering


@pytest.mark.parametrize("ar", ARITHS, ids=lambda a: a.name)
@pytest.mark.parametrize("entering,leaving", G.GET_LEAVING["cases"])
def test_get_leaving(ar, entering, leaving):
    st = LPState(_conv2(ar, G.GET_LEAVING["A"]), _conv(ar, G.GET_LEAVING["b"]), [], 4, 4, arith=ar)
    assert st.get_leaving(entering) == leaving


def _check_pivot(ar, vec, case):
    m, n = vec["m"], vec["n"]
    variables, coefficients = _names(m + n)
    st = LPState(_conv2(ar, vec["A"]), _conv(ar, vec["b"]), _conv(ar, vec["c"]), m, n,
                 variables=variables, coefficients=coefficients, arith=ar)
    st.pivot(case["e"], case["l"])
    for got, want in zip(st.A, case["resA"]):
        assert _eq(ar, got, want)
    assert _eq(ar, st.b, case["resB"])
    assert _eq(ar, st.c, case["resC"])
    assert ar.cmp(st.v, ar.conv(case["resV"])) == 0
    assert st.variables == case["resVariables"]
    assert st.coefficients == case["resCoefficients"]


@pytest.mark.parametrize("ar", ARITHS, ids=lambda a: a.name)
def test_pivot_1x1(ar):
    _check_pivot(ar, G.PIVOT_1x1, G.PIVOT_1x1)


@pytest.mark.parametrize("ar", ARITHS, ids=lambda a: a.name)
@pytest.mark.parametrize("vec", [G.PIVOT_2x2, G.PIVOT_4x5, G.PIVOT_7x2], ids=["2x2", "4x5", "7x2"])
def test_pivot_vectors(ar, vec):
    for case in vec["cases"]:
        _check_pivot(ar, vec, case)


@pytest.mark.parametrize("ar", ARITHS, ids=lambda a: a.name)
@pytest.mark.parametrize("b,answer", G.MIN_IN_B)
def test_min_in_b(ar, b, answer):
    assert LPSolver(ar).min_in_b(_conv(ar, b)) == answer


@pytest.mark.parametrize("coefs", G.X0_NAMES)
def test_x0_name(coefs):
    assert LPSolver.get_name_for_x0(coefs) not in coefs


@pytest.mark.parametrize("ar", ARITHS, ids=lambda a: a.name)
def test_aux_construction(ar):
    v = G.AUX_CONSTRUCTION
    variables, coefficients = _names(v["n"])
    form = LPStandardForm(v["A"], v["b"], v["c"], v["m"], v["n"], True, variables, coefficients, arith=ar)
    st = LPSolver(ar).convert_into_aux_lp(form)
    assert len(st.coefficients) == len(st.variables)
    for got, want in zip(st.A, v["resA"]):
        assert _eq(ar, got, want)
    assert _eq(ar, st.c, v["resC"])
    assert "x0" in st.coefficients and "x0" in st.variables.values()


def test_slack_form_names():
    # LPSolverSpec.groovy:59-74
    variables = {0: "x0", 1: "x1", 2: "x4", 3: "x6"}
    coefficients = {"x0": 0, "x1": 1, "x4": 2, "x6": 3}
    form = LPStandardForm([[0] * 4] * 4, [0] * 4, [0] * 4, 4, 4, True, variables, coefficients)
    st = LPSolver().convert_into_slack_form(form)
    assert len(st.variables) == 8 and len(st.coefficients) == 8


def _solve_case(ar, case, fix=False):
    names = case["names"]
    variables = coefficients = None
    if names is not None:
        variables = {i: nm for i, nm in enumerate(names)}
        coefficients = {nm: i for i, nm in enumerate(names)}
    form = LPStandardForm(case["A"], case["b"], case["c"], case["m"], case["n"], case["maximize"],
                          variables, coefficients, arith=ar)
    solver = LPSolver(ar, fix_restore_index=fix)
    verdict, value, message = "optimal", None, None
    try:
        value = solver.solve(form)
    except SolutionException as ex:
        verdict, message = "unbounded", str(ex)
    except LPException as ex:
        verdict, message = "infeasible", str(ex)
    return solver, verdict, value, message


@pytest.mark.parametrize("ar", ARITHS, ids=lambda a: a.name)
@pytest.mark.parametrize("case", G.SOLVE, ids=lambda c: c["name"])
def test_solve_known_answers(ar, case):
    solver, verdict, value, message = _solve_case(ar, case)
    assert verdict == case["verdict"]
    if verdict == "optimal":
        assert str(value) == case["value"]
    else:
        assert message == case["message"]
    tr = solver.trace
    if case.get("aux_log") is not None:
        assert tr.aux_state.name_log == case["aux_log"]
        assert tr.x0_final_index == case["x0_index"]
    if case.get("log") is not None:
        assert tr.final_state.name_log == case["log"]


def test_minimization_raw_v_is_not_17():
    # in decimal-15 the raw phase-2 v is 16.9999999999999; only setScale(6) makes it 17
    case = [c for c in G.SOLVE if c["name"] == "minimization"][0]
    solver, _, value, _ = _solve_case(Dec15, case)
    assert str(solver.trace.raw_v) == "16.9999999999999"
    assert str(value) == "-17.000000"


@pytest.mark.parametrize("ar", ARITHS, ids=lambda a: a.name)
def test_aux_lp_solving(ar):
    v = G.AUX_SOLVE
    variables = dict(v["variables"])
    coefficients = {nm: i for i, nm in variables.items()}
    aux = LPState(_conv2(ar, v["A"]), _conv(ar, v["b"]), _conv(ar, v["c"]), v["m"], v["n"],
                  variables=variables, coefficients=coefficients, arith=ar)
    x0 = LPSolver(ar).solve_aux_lp(aux, v["index_of_x0"], v["min_in_b"])
    assert ar.cmp(aux.v, ar.conv(v["resV"])) == 0
    assert x0 == v["x0_index"]
    assert aux.name_log == v["log"]


@pytest.mark.parametrize("ar", ARITHS, ids=lambda a: a.name)
@pytest.mark.parametrize("order", ["java", "insertion", "index"])
def test_restore_initial_lp(ar, order):
    v = G.RESTORE
    variables = dict(v["variables"])
    coefficients = {nm: i for i, nm in variables.items()}
    aux = LPState(_conv2(ar, v["A"]), _conv(ar, v["b"]), _conv(ar, v["c"]), v["m"], v["n"],
                  v=ar.ZERO, variables=variables, coefficients=coefficients, arith=ar)
    iv = dict(v["init_variables"])
    initial = LPStandardForm([[]], [], v["init_c"], v["init_m"], v["init_n"], True, iv,
                             {nm: i for i, nm in iv.items()}, arith=ar)
    initial.key_order = order
    res = LPSolver(ar).restore_initial_lp(aux, initial, v["index_of_x0"])
    for got, want in zip(res.A, v["resA"]):
        assert _eq(ar, got, want)
    assert _eq(ar, res.b, v["resB"]) and _eq(ar, res.c, v["resC"])
    assert ar.cmp(res.v, ar.conv(v["resV"])) == 0
    assert res.variables == v["resVariables"] and res.coefficients == v["resCoefficients"]


# ---- io_files/input.txt ------------------------------------------------------------------
with open(os.path.join(HERE, "golden", "input_txt_lps.json")) as _f:
    INPUT_LPS = json.load(_f)["lps"]


@pytest.mark.parametrize("ar", ARITHS, ids=lambda a: a.name)
def test_input_txt_lp1_matches_output_txt(ar):
    # readLP(File) stops at the first blank line => LP #1; io_files/output.txt:214-233
    text = "\n\n".join(e["text"] for e in INPUT_LPS)
    form = LPInputReader(ar).read_lp_file_text(text)
    assert (form.m, form.n) == (14, 18)
    solver = LPSolver(ar)
    assert str(solver.solve(form)) == G.INPUT_TXT_LP1["value"]
    x = primal_solution(solver.trace.final_state, 18)
    assert _eq(ar, x, G.INPUT_TXT_LP1["primal"])
    # current-rule sequence: 16 pivots, not the stale 8 of output.txt
    assert solver.trace.phase2_log == [(0, 0), (2, 7), (1, 0), (3, 1), (5, 2), (8, 9), (1, 8), (6, 0),
                                       (9, 2), (11, 10), (7, 0), (12, 4), (14, 5), (17, 12), (7, 11),
                                       (16, 0)]


@pytest.mark.parametrize("entry", INPUT_LPS, ids=lambda e: "lp%d" % e["index"])
@pytest.mark.parametrize("fix", [False, True], ids=["asref", "fixed"])
def test_input_txt_regression(entry, fix):
    """The committed golden JSON reproduces (guards the oracle against drift) and the two
    number systems agree on verdict and 6-decimal objective wherever the restore defect of
    LPSolver.java:220,231 is not in play."""
    for ar in ARITHS:
        want = entry["%s_%s" % (ar.name, "fixed" if fix else "asref")]
        if want["verdict"] == "parse_error":
            with pytest.raises(LPException):
                LPInputReader(ar).read_lp(entry["text"])
            continue
        form = LPInputReader(ar).read_lp(entry["text"])
        solver = LPSolver(ar, fix_restore_index=fix)
        try:
            val = solver.solve(form)
            assert want["verdict"] == "optimal" and str(val) == want["value"]
        except SolutionException as ex:
            assert want["verdict"] == "unbounded" and str(ex) == want["message"]
        except LPException as ex:
            assert want["verdict"] == "infeasible" and str(ex) == want["message"]
        assert [list(p) for p in solver.trace.phase1_log] == want["phase1_log"]
        assert [list(p) for p in solver.trace.phase2_log] == want["phase2_log"]
    if fix:
        d, f = entry["dec15_fixed"], entry["f64_fixed"]
        assert d["verdict"] == f["verdict"] and d.get("value") == f.get("value")


def test_input_txt_fixed_values_are_true_optima():
    # independent check (HiGHS, run when the fixture was generated): the index-shifted restore
    # gives the true optimum on every parsable LP of io_files/input.txt
    want = {1: "7.000000", 3: "10.000000", 4: "-10.000000", 5: "159.000000", 6: "138.333333",
            7: "3.000000", 8: "-20.000000", 9: "20.545455", 10: "6.000000", 12: "-20.000000",
            13: "-18.666667", 15: "9.000000"}
    for e in INPUT_LPS:
        if e["index"] in want:
            assert e["dec15_fixed"]["value"] == want[e["index"]]
    assert INPUT_LPS[13]["dec15_fixed"]["verdict"] == "infeasible"
