"""linear_programming_solver_b200 — a B200-native (sm_100a) dense-tableau simplex pivot loop
behind the API of Toptachamann/Linear_Programming_Solver's `lpsolver` package.

The product is `liblps_b200.so` (hand-written CUDA kernels + a C ABI, `include/lps_b200.h`,
and a C++ host driver, `include/lpsolver_host.h`); this package is its ctypes face with the
reference's class names.  There is no CPU fallback.
"""
from .exceptions import LPException, LpsError, SolutionException  # noqa: F401
from .lp_input_reader import LPInputReader  # noqa: F401
from .lp_solver import LPSolver  # noqa: F401
from .lp_standard_form import LPStandardForm  # noqa: F401
from .lp_state import LPState  # noqa: F401

__all__ = ["LPSolver", "LPState", "LPStandardForm", "LPInputReader", "LPException", "SolutionException", "LpsError"]
