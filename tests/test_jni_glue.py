"""The Java side of the drop-in (integration/java) cannot be compiled here (no JDK), but its C half can:
integration/java/jni/lps_b200_jni.c is built against a stand-in <jni.h> (tests/jni_mock/jni.h) and

  * (CPU)  must export exactly one Java_lpsolver_LPStateNative_<name> per `native` method that
           integration/java/lpsolver/LPStateNative.java declares, with matching parameter counts;
  * (GPU)  is driven by tests/jni_mock/jni_call_order.c in the order LPStateNative / the patched LPSolver call
           it (borrowed primitive arrays, every borrow released before the native function returns) and must
           reproduce the binary64 oracle: verdict, objective, pivot log, b, c, positions, primal.
"""
import json
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JAVA = os.path.join(ROOT, "integration", "java", "lpsolver", "LPStateNative.java")
GLUE = os.path.join(ROOT, "integration", "java", "jni", "lps_b200_jni.c")
MOCK = os.path.join(ROOT, "tests", "jni_mock")
LIBDIR = os.path.join(ROOT, "linear_programming_solver_b200")


def _build(tmp_path):
    glue_o = str(tmp_path / "glue.o")
    drv_o = str(tmp_path / "drv.o")
    exe = str(tmp_path / "jni_call_order")
    inc = ["-I" + MOCK, "-I" + os.path.join(ROOT, "include")]
    subprocess.check_call(["gcc", "-O1", "-Wall", "-Wextra", "-Werror"] + inc + ["-c", GLUE, "-o", glue_o])
    subprocess.check_call(["gcc", "-O1", "-Wall", "-Wextra"] + inc + ["-c", os.path.join(MOCK, "jni_call_order.c"), "-o", drv_o])
    subprocess.check_call(["gcc", glue_o, drv_o, "-L" + LIBDIR, "-llps_b200", "-lm", "-Wl,-rpath," + LIBDIR, "-o", exe])
    return glue_o, exe


def _java_natives():
    src = open(JAVA).read()
    out = {}
    for m in re.finditer(r"private static native\s+\S+\s+(\w+)\(([^)]*)\);", src):
        args = [a for a in m.group(2).split(",") if a.strip()]
        out[m.group(1)] = len(args)
    return out


def test_glue_compiles_and_matches_the_java_native_declarations(tmp_path):
    glue_o, _ = _build(tmp_path)
    syms = subprocess.check_output(["nm", "--defined-only", glue_o], text=True)
    exported = set(re.findall(r" T Java_lpsolver_LPStateNative_(\w+)", syms))
    natives = _java_natives()
    assert len(natives) >= 20
    assert exported == set(natives), (sorted(exported - set(natives)), sorted(set(natives) - exported))
    # parameter counts: JNIEnv*, jclass + the Java parameters
    csrc = open(GLUE).read()
    for name, nargs in natives.items():
        if name in ("nReadB", "nReadC"):          # instances of the READ_VECTOR macro: (env, cls, h, out)
            assert "READ_VECTOR(%s," % name in csrc and nargs == 2
            continue
        m = re.search(r"FN\(%s\)\(([^)]*)\)" % name, csrc)
        assert m, name
        assert len([a for a in m.group(1).split(",") if a.strip()]) == nargs + 2, name


def _write_lp(path, A, b, c, maximize=True):
    m, n = A.shape
    with open(path, "w") as f:
        f.write("%d %d %d\n" % (m, n, 1 if maximize else 0))
        for row in A:
            f.write(" ".join(repr(float(x)) for x in row) + "\n")
        f.write(" ".join(repr(float(x)) for x in b) + "\n")
        f.write(" ".join(repr(float(x)) for x in c) + "\n")


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["run", "loop"])
def test_jni_call_order_reproduces_the_oracle(tmp_path, mode):
    from oracle import tier_f
    _, exe = _build(tmp_path)
    m, n = 37, 53
    A, b, c = tier_f.gen_dense_feasible(m, n, 3)
    ref = tier_f.TierFState(A.copy(), b.copy(), c.copy())
    status, k = ref.run()
    lp = str(tmp_path / "lp.txt")
    _write_lp(lp, A, b, c)
    out = json.loads(subprocess.check_output([exe, lp, mode], text=True))
    assert out["verdict"] == 1 and status == tier_f.OPTIMAL and out["pivots"] == k
    assert [tuple(out["log"][2 * i:2 * i + 2]) for i in range(k)] == [tuple(x) for x in ref.log]
    assert out["v"] == ref.v[0]
    assert np.array_equal(np.array(out["b"]), ref.b) and np.array_equal(np.array(out["c"]), ref.c)
    assert np.array_equal(np.array(out["positions"]), ref.pos2var)


@pytest.mark.gpu
def test_jni_call_order_phase1(tmp_path):
    """LPSolverSpec.groovy:100-111 (infeasible start -> 20) and a random LP with negative right-hand sides,
    through nLoadAux / nPivot / nRun / nPositionOf / nDropColumn / nRebuildObjective"""
    import linear_programming_solver_b200 as L
    _, exe = _build(tmp_path)
    A = np.array([[1, 0], [-1, 0], [0, 1], [0, -1]], dtype=np.float64)
    b = np.array([10, -2, 10, -2], dtype=np.float64)
    c = np.array([1, 1], dtype=np.float64)
    lp = str(tmp_path / "lp1.txt")
    _write_lp(lp, A, b, c)
    out = json.loads(subprocess.check_output([exe, lp, "phase1"], text=True))
    assert out["verdict"] == 1 and abs(out["v"] - 20.0) < 1e-12
    rng = np.random.default_rng(11)
    m, n = 14, 9
    A = rng.integers(-4, 9, size=(m, n)).astype(np.float64)
    x = rng.integers(0, 4, size=n).astype(np.float64)
    b = A @ x + rng.integers(0, 5, size=m)
    b[::3] -= 30.0
    c = rng.integers(-3, 6, size=n).astype(np.float64)
    lp = str(tmp_path / "lp2.txt")
    _write_lp(lp, A, b, c)
    out = json.loads(subprocess.check_output([exe, lp, "phase1"], text=True))
    solver = L.LPSolver(fix_restore_index=True)
    try:
        want = solver.solve(L.LPStandardForm(A, b, c, m, n, True))
        assert out["verdict"] == 1 and abs(out["v"] - float(want)) <= 1e-6
    except L.SolutionException as ex:
        assert out["verdict"] in (2, str(ex))
    except L.LPException as ex:
        assert out["verdict"] == str(ex)
