"""The native LPInputReader (csrc/lp_input_reader.cpp) against the reference's parser tests
(LPInputReaderSpec.groovy, LPInputReaderTest.java) and, by fuzzing, against the oracle's
regex-based restatement of LPInputReader.java.  CPU only (host code, no device)."""
import json
import os
import random

import numpy as np
import pytest

import linear_programming_solver_b200 as L
from oracle.arith import F64
from oracle.lp_text import LPInputReader as OracleReader
from oracle.simplex_ref import LPException as OLPException

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "input_txt_lps.json")) as _f:
    INPUT_LPS = json.load(_f)["lps"]


def test_simple_lp_reading():
    # LPInputReaderSpec.groovy:7-25
    f = L.LPInputReader().read_lp("max\nx1 + x2\nx1 + x2 <= 0")
    assert f.A.tolist() == [[1.0, 1.0]] and f.b.tolist() == [0.0] and f.c.tolist() == [1.0, 1.0]
    assert f.variables == {0: "x1", 1: "x2"} and f.coefficients == {"x1": 0, "x2": 1}
    assert (f.m, f.n, f.maximize) == (1, 2, True)


def test_complicated_lp_reading():
    # LPInputReaderSpec.groovy:27-50 (40-digit literals land on the nearest binary64)
    lp = ("min\n"
          "        782343246437439743943794343944324324324*x1  + 5273392392323.238324379948439874973439732242x2       \n"
          "          6.338203729*x1   +   0.732932323x2 >=  9    \n"
          "     -   102333.233232x1 + 2332.33214*x2 ==   13    \n"
          "      11x1  -  x2   =  -   5435377467645646394439874397439347934734   ")
    f = L.LPInputReader().read_lp(lp)
    assert f.A.tolist() == [[-6.338203729, -0.732932323], [-102333.233232, 2332.33214], [102333.233232, -2332.33214],
                            [11.0, -1.0], [-11.0, 1.0]]
    assert f.b.tolist() == [-9.0, 13.0, -13.0, -5435377467645646394439874397439347934734.0,
                            5435377467645646394439874397439347934734.0]
    assert f.c.tolist() == [782343246437439743943794343944324324324.0, 5273392392323.238324379948439874973439732242]
    assert f.variables == {0: "x1", 1: "x2"} and (f.m, f.n, f.maximize) == (5, 2, False)


def test_max_min_parameter():
    # LPInputReaderTest.java:28-49
    r = L.LPInputReader()
    for s in ("max", "MAX", "mAx", "  Max  "):
        assert r.read_lp(s + "\nx1\nx1 <= 1").maximize
    for s in ("min", "MIN", "mIn"):
        assert not r.read_lp(s + "\nx1\nx1 <= 1").maximize
    with pytest.raises(L.LPException, match="Incorrect max/min parameter"):
        r.read_lp("maximize\nx1\nx1 <= 1")


def test_objective_processing():
    # LPInputReaderTest.java:51-96
    r = L.LPInputReader()
    assert r.read_lp("max\nx1\nx1 <= 1").c.tolist() == [1.0]
    f = r.read_lp("max\n5*x1 + 1.23x3 + 5.32*vv\nx1 <= 1")
    assert f.c.tolist() == [5.0, 1.23, 5.32] and f.variables == {0: "x1", 1: "x3", 2: "vv"}
    with pytest.raises(L.LPException, match="Can't recognize objective"):
        r.read_lp("max\n\nx1 <= 1")
    with pytest.raises(ValueError):
        r.read_lp(None)


def test_constraint_errors_and_incomplete_lps():
    # LPInputReaderTest.java:98-163
    r = L.LPInputReader()
    for bad in (" 5*x1 + 1.23x3 + 5.32*vv ", " 5*x1 + 1.23x3 + 5.32*vv <=  ", ""):
        with pytest.raises(L.LPException, match="Can't recognize constraint"):
            r.read_lp("max\nx1\nx1 <= 1\n" + bad + "\nx1 <= 2")
    for bad in ("", "x1 + x2\nx1 + x2 <= 0", "max\nx1 + x2 <= 0", "max\nx1 + x2"):
        with pytest.raises(L.LPException):
            r.read_lp(bad)


def test_new_variable_in_constraint_and_lowering():
    # README example: variables first seen in a constraint extend c with 0 (LPInputReader.java:172-178)
    f = L.LPInputReader().read_lp("max\n2x1 + 3.05*x3\n1.05*x4 + 25*x1 == 0\n3.66x1 = 3\n2x2 + x3 <= 0\nx1 + x2 + x3 + x24 >= 0")
    assert f.variables == {0: "x1", 1: "x3", 2: "x4", 3: "x2", 4: "x24"}
    assert f.c.tolist() == [2.0, 3.05, 0.0, 0.0, 0.0]
    assert (f.m, f.n) == (6, 5)
    assert f.A.tolist() == [[25, 0, 1.05, 0, 0], [-25, 0, -1.05, 0, 0], [3.66, 0, 0, 0, 0], [-3.66, 0, 0, 0, 0],
                            [0, 1, 0, 2, 0], [-1, -1, 0, -1, -1]]
    assert f.b.tolist() == [0, 0, 3, -3, 0, 0]


@pytest.mark.parametrize("entry", INPUT_LPS, ids=lambda e: "lp%d" % e["index"])
def test_input_txt_equals_oracle_parser(entry):
    try:
        want = OracleReader(F64).read_lp(entry["text"])
    except OLPException as ex:
        with pytest.raises(L.LPException, match=str(ex).replace("'", ".")):
            L.LPInputReader().read_lp(entry["text"])
        return
    got = L.LPInputReader().read_lp(entry["text"])
    assert (got.m, got.n, got.maximize) == (want.m, want.n, want.maximize)
    assert got.A.tolist() == want.A and got.b.tolist() == want.b and got.c.tolist() == want.c
    assert got.variables == want.variables and got.coefficients == want.coefficients


def test_file_semantics_stop_at_blank_line(tmp_path):
    text = "\n\n".join(e["text"] for e in INPUT_LPS)
    p = tmp_path / "input.txt"
    p.write_text(text)
    f = L.LPInputReader().read_lp_file(str(p))
    assert (f.m, f.n) == (14, 18)                      # LP #1 only (LPInputReader.java:80-81)
    with pytest.raises(ValueError):
        L.LPInputReader().read_lp_file(str(tmp_path / "missing.txt"))
    empty = tmp_path / "empty.txt"
    empty.write_text("")
    with pytest.raises(L.LPException, match="Input file is empty"):
        L.LPInputReader().read_lp_file(str(empty))


def _random_line(rng, constraint=True):
    wild = rng.random() < 0.25              # a quarter of the lines use the junk alphabet too
    pieces = []
    for k in range(rng.randrange(1, 5)):
        sign = rng.choice(["+", "-", " + ", " - ", "+ ", " -"] + ([""] if (k == 0 or wild) else []))
        num = rng.choice(["", "2", "3.5", ".5", "7.", "12", "0.25"] + (["."] if wild else []))
        star = rng.choice(["", "*", ""] + ([" * ", " "] if wild else []))
        name = rng.choice(["x1", "x2", "y", "var12", "X", "x"] + (["1", "x1x2", ""] if wild else []))
        pieces.append(sign + num + star + name)
    line = "".join(pieces)
    if not constraint:
        return line + rng.choice(["", " ", "  "])
    tails = [" <= 4", ">= 2.5", " = -3", "== - 1", " = 7 ", "<=0", " >= 12.25  "]
    if wild:
        tails += [" <= ", " < 3", "<= 3 4", " => 2", "<= 2.", "<= -.5", ""]
    return line + rng.choice(tails)


def test_fuzz_against_oracle_regex_restatement():
    rng = random.Random(2024)
    n_ok = n_bad = 0
    for _ in range(4000):
        lines = [rng.choice(["max", "min", "Max", " MIN "]), _random_line(rng, constraint=False)]
        lines += [_random_line(rng) for _ in range(rng.randrange(1, 4))]
        text = "\n".join(lines)
        try:
            want = OracleReader(F64).read_lp(text)
            werr = None
        except OLPException as ex:
            want, werr = None, str(ex)
        except Exception:                       # decimal.InvalidOperation <-> NumberFormatException
            want, werr = None, "number"
        try:
            got = L.LPInputReader().read_lp(text)
            gerr = None
        except L.LPException as ex:
            got, gerr = None, str(ex)
        except ValueError:
            got, gerr = None, "number"
        assert werr == gerr, (text, werr, gerr)
        if want is not None:
            n_ok += 1
            assert got.A.tolist() == want.A and got.b.tolist() == want.b and got.c.tolist() == want.c, text
            assert got.variables == want.variables and got.coefficients == want.coefficients, text
        else:
            n_bad += 1
    assert n_ok > 300 and n_bad > 300


def test_get_dual():
    # LPStandardFormSpec.groovy:6-25
    names = {0: "x1", 1: "x2", 2: "x3", 3: "x4"}
    form = L.LPStandardForm([[1, -2, -1, 3], [3, 1, 0, 4], [3, 4, 2, 2]], [3, 4, 1], [4, 1, 2, 3], 3, 4, True, names,
                            {v: k for k, v in names.items()})
    dual = form.get_dual()
    assert dual.A.tolist() == [[1, 3, 3], [-2, 1, 4], [-1, 0, 2], [3, 4, 2]]
    assert dual.b.tolist() == [4, 1, 2, 3] and dual.c.tolist() == [3, 4, 1]
    assert (dual.m, dual.n, dual.maximize) == (4, 3, False)
