"""ctypes binding of liblps_b200.so (include/lps_b200.h, include/lpsolver_host.h).

The shared library is the product; this module only declares its C ABI.  There is no CPU
fallback: if the library is missing or no CUDA device is present, loading or `lps_create`
fails loudly.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char, c_char_p, c_double, c_float, c_int, c_int64,
                    c_uint64, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblps_b200.so")

LPS_OK = 0
LPS_ERR_INVALID, LPS_ERR_CUDA, LPS_ERR_STATE, LPS_ERR_NOMEM, LPS_ERR_NODEVICE, LPS_ERR_COMM = -1, -2, -3, -4, -5, -6
LPS_RUNNING, LPS_OPTIMAL, LPS_UNBOUNDED, LPS_PIVOT_CAP = 0, 1, 2, 3
LPS_GEN_DENSE, LPS_GEN_UNBOUNDED, LPS_GEN_ASSIGNMENT = 0, 1, 2
(LPSOLVER_OPTIMAL, LPSOLVER_UNBOUNDED, LPSOLVER_INFEASIBLE, LPSOLVER_AUX_UNBOUNDED,
 LPSOLVER_DEGENERATE_FAIL, LPSOLVER_INDEX_ERROR, LPSOLVER_PIVOT_CAP, LPSOLVER_ERROR) = range(8)


class LpsOptions(Structure):
    _fields_ = [("epsilon", c_double), ("inf", c_double), ("device", c_int), ("time_kernels", c_int),
                ("stream", c_void_p), ("update_variant", c_int), ("loop_mode", c_int), ("block_pivots", c_int),
                ("panel_ctas", c_int), ("pass_chunk_rows", c_int), ("reserved", c_int * 3)]


class LpsRunResult(Structure):
    _fields_ = [("verdict", c_int), ("last_entering", c_int), ("last_leaving", c_int), ("pad_", c_int),
                ("npivots", c_int64), ("total_pivots", c_int64), ("v", c_double),
                ("device_ms", c_float), ("update_ms", c_float), ("update_launches", c_int64),
                ("kernel_launches", c_int64)]


class LpsObjectiveOp(Structure):
    _fields_ = [("kind", c_int), ("index", c_int), ("coef", c_double)]


class LpsolverResult(Structure):
    _fields_ = [("verdict", c_int), ("used_phase1", c_int), ("x0_index", c_int), ("pad_", c_int),
                ("phase1_pivots", c_int64), ("phase2_pivots", c_int64), ("value", c_double),
                ("device_ms", c_float), ("pad2_", c_float), ("value6", c_char * 48),
                ("message", c_char * 160)]


_dp = POINTER(c_double)
_ip = POINTER(c_int)

# every symbol include/lps_b200.h and include/lpsolver_host.h declare: (restype, argtypes)
SIGNATURES = {
    "lps_abi_version": (c_int, []),
    "lps_default_options": (None, [POINTER(LpsOptions)]),
    "lps_status_string": (c_char_p, [c_int]),
    "lps_create": (c_int, [POINTER(c_void_p), POINTER(LpsOptions)]),
    "lps_destroy": (c_int, [c_void_p]),
    "lps_last_error": (c_char_p, [c_void_p]),
    "lps_load": (c_int, [c_void_p, c_int, c_int, _dp, c_int64, _dp, _dp, c_double]),
    "lps_load_aux": (c_int, [c_void_p, c_int, c_int, _dp, c_int64, _dp]),
    "lps_generate_dense": (c_int, [c_void_p, c_int, c_int, c_uint64, c_int]),
    "lps_generate_lp": (c_int, [c_void_p, c_int, c_int, c_int, c_uint64, c_int]),
    "lps_get_entering": (c_int, [c_void_p, _ip]),
    "lps_get_leaving": (c_int, [c_void_p, c_int, _ip]),
    "lps_pivot": (c_int, [c_void_p, c_int, c_int]),
    "lps_run": (c_int, [c_void_p, c_int64, POINTER(LpsRunResult)]),
    "lps_dims": (c_int, [c_void_p, _ip, _ip]),
    "lps_read_v": (c_int, [c_void_p, _dp]),
    "lps_read_b": (c_int, [c_void_p, _dp]),
    "lps_read_c": (c_int, [c_void_p, _dp]),
    "lps_read_row": (c_int, [c_void_p, c_int, _dp]),
    "lps_read_col": (c_int, [c_void_p, c_int, _dp]),
    "lps_read_A": (c_int, [c_void_p, _dp, c_int64]),
    "lps_read_positions": (c_int, [c_void_p, _ip]),
    "lps_position_of": (c_int, [c_void_p, c_int, _ip]),
    "lps_read_pivot_log": (c_int, [c_void_p, _ip, c_int64, POINTER(c_int64)]),
    "lps_read_primal": (c_int, [c_void_p, c_int, _dp]),
    "lps_first_nonzero_in_row": (c_int, [c_void_p, c_int, _ip]),
    "lps_drop_column": (c_int, [c_void_p, c_int]),
    "lps_rebuild_objective": (c_int, [c_void_p, POINTER(LpsObjectiveOp), c_int]),
    "lps_device_info": (c_int, [c_void_p, _ip, POINTER(c_int64), _ip, _ip]),
    "lps_tableau_bytes": (c_int, [c_void_p, POINTER(c_int64)]),
    "lps_algorithmic_bytes_per_pivot": (c_int, [c_void_p, POINTER(c_int64)]),
    "lps_measure_fp64_issue_rate": (c_int, [c_void_p, c_double, _dp]),
    "lps_loop_description": (c_int, [c_void_p, c_char_p, c_int]),
    "lps_plan_split_model": (c_int, [c_int, c_int, c_int, c_int64, c_int64]),
    "lps_plan_split_tuned": (c_int, [c_int, c_int, c_int, c_int, c_double, c_double]),
    "lps_shard_generate_dense": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_uint64, c_int]),
    "lps_shard_generate_lp": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_uint64, c_int]),
    "lps_shard_load": (c_int, [c_void_p, c_int, c_int, c_int, c_int, _dp, c_int64, _dp, _dp, c_double]),
    "lps_shard_info": (c_int, [c_void_p, _ip, _ip, _ip, _ip, _ip]),
    "lps_shard_export": (c_int, [c_void_p, c_void_p]),
    "lps_shard_comm_ptr": (c_int, [c_void_p, POINTER(c_void_p)]),
    "lps_shard_attach_ipc": (c_int, [c_void_p, c_void_p]),
    "lps_shard_attach_ptrs": (c_int, [c_void_p, POINTER(c_void_p)]),
    "lpsolver_solve": (c_int, [POINTER(LpsOptions), c_int, c_int, _dp, c_int64, _dp, _dp, c_int, c_int,
                               c_int64, POINTER(LpsolverResult), _dp, _ip, c_int64, _ip, c_int64,
                               POINTER(c_void_p)]),
    "lpsolver_read_lp": (c_int, [c_char_p, c_int, _ip, _ip, _ip, POINTER(_dp), POINTER(_dp), POINTER(_dp),
                                 POINTER(c_void_p), c_char_p, c_int]),
    "lpsolver_free": (None, [c_void_p]),
    "lpsolver_set_scale6": (c_int, [c_double, c_char_p, c_int]),
    "lpsolver_min_in_b": (c_int, [_dp, c_int]),
}

_LIB = None


def load():
    """Load liblps_b200.so and declare its signatures.  Raises if the library is not built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "liblps_b200.so is not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C linear_programming_solver_b200/csrc`.  There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header/library mismatch
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.lps_abi_version() != 1:
            raise ImportError("liblps_b200.so ABI version mismatch")
        _LIB = lib
    return _LIB


def default_options() -> LpsOptions:
    o = LpsOptions()
    load().lps_default_options(byref(o))
    return o
