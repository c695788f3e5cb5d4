"""The C-ABI library loads and exports every symbol include/*.h declares (no GPU needed), and the
host-only helpers work.  Also guards the product/oracle separation."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(path):
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(lps_[A-Za-z0-9_]+|lpsolver_[A-Za-z0-9_]+)\s*\(", text))
    return {n for n in names if not n.endswith("_s")}


def test_every_declared_symbol_is_exported_and_bound():
    from linear_programming_solver_b200 import _native as N
    lib = N.load()
    declared = _declared(os.path.join(ROOT, "include", "lps_b200.h")) | _declared(
        os.path.join(ROOT, "include", "lpsolver_host.h"))
    declared -= {"lps_handle", "lps_status", "lps_verdict", "lps_options", "lps_run_result",
                 "lps_objective_op", "lpsolver_result"}
    assert len(declared) >= 35
    for name in sorted(declared):
        assert hasattr(lib, name), "liblps_b200.so does not export %s" % name
        assert name in N.SIGNATURES, "ctypes binding lacks %s" % name
    assert set(N.SIGNATURES) <= declared, set(N.SIGNATURES) - declared
    assert lib.lps_abi_version() == 1


def test_struct_layouts_match_header_sizes():
    from linear_programming_solver_b200 import _native as N
    assert ctypes.sizeof(N.LpsOptions) == 64       # static_asserts in csrc/lps_api.cu pin the C side
    assert ctypes.sizeof(N.LpsObjectiveOp) == 16
    assert ctypes.sizeof(N.LpsRunResult) == 64
    assert ctypes.sizeof(N.LpsolverResult) == 16 + 16 + 8 + 8 + 48 + 160


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import linear_programming_solver_b200 as L
    with pytest.raises(L.LpsError) as ei:
        L.LPState([[1.0]], [1.0], [1.0], 1, 1)
    assert "no CUDA device" in str(ei.value)
    with pytest.raises(L.LpsError):
        L.LPSolver().solve(L.LPStandardForm([[1.0]], [1.0], [1.0], 1, 1, True))


def test_set_scale6_and_min_in_b_host_helpers():
    import decimal
    from decimal import Decimal

    from linear_programming_solver_b200 import _native as N
    lib = N.load()
    decimal.getcontext().prec = 400

    def s6(v):
        buf = ctypes.create_string_buffer(2000)
        lib.lpsolver_set_scale6(v, buf, 2000)
        return buf.value.decode()

    for v in [0.0, -0.0, 1.0, 7.999999999999999, 16.9999999999999, -17.0, 0.0000005, 0.00000049999999999,
              -0.0000005, 123456.7890125, 2.5e-7, 1e22, -3.3333335, 20.545454545454547, 0.9999995, 138.33333333333331]:
        want = str(Decimal(v).quantize(Decimal("0.000001"), rounding=decimal.ROUND_HALF_UP))
        if want.startswith("-") and float(want) == 0:
            want = want[1:]
        assert s6(v) == want
    arr = (ctypes.c_double * 4)(-1, -1000, -10, -1001)
    assert lib.lpsolver_min_in_b(arr, 4) == 3          # LPSolverSpec.groovy:19


def test_product_does_not_import_oracle():
    """The product path must never route through the CPU oracle."""
    pkg = os.path.join(ROOT, "linear_programming_solver_b200")
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|from\s+\.\.?oracle)|#include\s+[\"<].*oracle|dlopen|CDLL\(.*tier", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), "%s references the oracle" % f
