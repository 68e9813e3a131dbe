"""ctypes binding of libtensorkrylov_b200.so (include/tensorkrylov_b200.h).

This is the same set of symbols the Julia `ccall` wrapper binds
(julia/TensorKrylovB200.jl).  There is no CPU fallback: if the shared library
is missing the import fails, and every compute entry point fails with the
library's own error when no B200 is visible.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtensorkrylov_b200.so")
TABLES_PATH = os.path.join(_HERE, "data", "expsum_tables.bin")

# enums of the header
TK_SYM, TK_NONSYM = 0, 1
TK_LAPLACE_DENSE, TK_LAPLACE, TK_CONVDIFF, TK_EIGVALMAT, TK_RANDSPD, TK_GENERIC = 0, 1, 2, 3, 4, 5
TK_LANCZOS, TK_LANCZOS_REORTH, TK_ARNOLDI = 0, 1, 2
TK_CONVERGED, TK_NMAX, TK_BREAKDOWN, TK_NAN, TK_RUNNING = 0, 1, 2, 3, -1
TK_FLAG_REFERENCE_H1, TK_FLAG_FIXED_ITERATIONS, TK_FLAG_TIME_KERNELS, TK_FLAG_TIME_ALL = 1, 2, 4, 8

EXPORTS = [
    "tk_last_error", "tk_version", "tk_device_count",
    "tk_tables_load", "tk_tables_sym_lookup", "tk_tables_sym_rank", "tk_nonsym_coefficients", "tk_laplace_extremes",
    "tk_comm_unique_id", "tk_create", "tk_destroy", "tk_release_cache", "tk_local_modes", "tk_needs_mode",
    "tk_set_operator_csc", "tk_set_operator_dense", "tk_share_operator", "tk_share_operator_all", "tk_set_rhs", "tk_set_rhs_all",
    "tk_set_schedule", "tk_schedule_laplace", "tk_schedule", "tk_minor_extremes", "tk_solve", "tk_solution_rank", "tk_get_solution", "tk_get_solution_all",
    "tk_get_solution_device", "tk_alloc_host", "tk_free_host", "tk_get_detail", "tk_get_solve_info",
    "tk_begin", "tk_step_bases", "tk_compress", "tk_residual",
    "tk_get_H", "tk_get_V", "tk_get_bt", "tk_get_Y", "tk_get_eig", "tk_get_orth_state",
    "tk_tridiag_eig_batched", "tk_timing_mark", "tk_get_timing", "tk_launch_count",
]


class TKError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libtensorkrylov_b200 error {code}: {msg}")
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C tensorkrylov.jl_b200/csrc`).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    p, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    pd = C.POINTER(C.c_double)
    pi32, pi64 = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    sig = {
        "tk_last_error": (C.c_char_p, []),
        "tk_version": (C.c_int, []),
        "tk_device_count": (C.c_int, [C.POINTER(C.c_int)]),
        "tk_tables_load": (C.c_int, [C.c_char_p]),
        "tk_tables_sym_lookup": (C.c_int, [f64, f64, pi32, pi32, pi32, pd, pd]),
        "tk_tables_sym_rank": (C.c_int, [f64, i32, pd, pd, pd]),
        "tk_nonsym_coefficients": (C.c_int, [f64, f64, i32, pi32, pi32, pd, pd]),
        "tk_laplace_extremes": (C.c_int, [i32, i64, i32, pd, pd]),
        "tk_comm_unique_id": (C.c_int, [p]),
        "tk_create": (C.c_int, [C.POINTER(p), i32, pi64, i32, i32, i32, i32, i32, i32, i32, i32, p]),
        "tk_destroy": (None, [p]),
        "tk_release_cache": (C.c_int, []),
        "tk_local_modes": (C.c_int, [p, pi32, pi32]),
        "tk_needs_mode": (C.c_int, [p, i32, pi32]),
        "tk_set_operator_csc": (C.c_int, [p, i32, i64, pi64, pi64, pd]),
        "tk_set_operator_dense": (C.c_int, [p, i32, i64, pd, C.c_char]),
        "tk_share_operator": (C.c_int, [p, i32, i32]),
        "tk_share_operator_all": (C.c_int, [p, i32]),
        "tk_set_rhs": (C.c_int, [p, i32, pd, i64]),
        "tk_set_rhs_all": (C.c_int, [p, pd, i64]),
        "tk_set_schedule": (C.c_int, [p, i32, f64, i32, pd, pd]),
        "tk_schedule_laplace": (C.c_int, [p, f64]),
        "tk_schedule": (C.c_int, [p, f64]),
        "tk_minor_extremes": (C.c_int, [pd, i32, i32, i32, pd]),
        "tk_solve": (C.c_int, [p, f64, pi32, pi64, pi32, pd, pd, pd]),
        "tk_solution_rank": (C.c_int, [p, pi32]),
        "tk_get_solution": (C.c_int, [p, i32, pd, i32, pd, i64, i32]),
        "tk_get_solution_all": (C.c_int, [p, pd, i32, pd, i64, i32]),
        "tk_get_solution_device": (C.c_int, [p, pd, i32, p, i64, i32]),
        "tk_alloc_host": (C.c_int, [C.POINTER(p), i64]),
        "tk_free_host": (C.c_int, [p]),
        "tk_get_detail": (C.c_int, [p, i32, i32, pd]),
        "tk_get_solve_info": (C.c_int, [p, pi32, pd, pi32, pi32]),
        "tk_begin": (C.c_int, [p]),
        "tk_step_bases": (C.c_int, [p, i32]),
        "tk_compress": (C.c_int, [p, i32]),
        "tk_residual": (C.c_int, [p, i32, f64, pd]),
        "tk_get_H": (C.c_int, [p, i32, pd]),
        "tk_get_V": (C.c_int, [p, i32, i32, pd]),
        "tk_get_bt": (C.c_int, [p, i32, pd]),
        "tk_get_Y": (C.c_int, [p, i32, i32, pd, pi32]),
        "tk_get_eig": (C.c_int, [p, i32, i32, pd, pd]),
        "tk_get_orth_state": (C.c_int, [p, i32, pd, pi32]),
        "tk_tridiag_eig_batched": (C.c_int, [i32, i32, i32, pd, pd, pd, pd, pi32]),
        "tk_timing_mark": (C.c_int, [p]),
        "tk_get_timing": (C.c_int, [p, i32, pd, pi64, pd]),
        "tk_launch_count": (C.c_int, [p, pi64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        raise TKError(rc, lib.tk_last_error().decode("utf-8", "replace"))


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


_tables_loaded = None


def load_tables(path=None):
    """tk_tables_load: the packed table file shipped with the package, or the
    reference's coefficients_data/ directory."""
    global _tables_loaded
    path = path or TABLES_PATH
    if _tables_loaded != path:
        check(lib.tk_tables_load(path.encode()))
        _tables_loaded = path


class PinnedArray:
    """A float64 numpy array over page-locked host memory from tk_alloc_host (results then cross PCIe by DMA,
    without the staging copy a pageable destination needs).  The memory is returned by close() or with the object."""

    def __init__(self, shape):
        self.shape = tuple(int(x) for x in shape)
        self.nbytes = 8 * int(np.prod(self.shape)) if self.shape else 8
        ptr = C.c_void_p()
        check(lib.tk_alloc_host(C.byref(ptr), self.nbytes))
        self.ptr = ptr
        buf = (C.c_double * max(self.nbytes // 8, 1)).from_address(ptr.value)
        self.array = np.frombuffer(buf, dtype=np.float64, count=int(np.prod(self.shape))).reshape(self.shape)

    def close(self):
        if getattr(self, "ptr", None):
            self.array = None
            lib.tk_free_host(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_count():
    n = C.c_int(0)
    rc = lib.tk_device_count(C.byref(n))
    return n.value if rc == 0 else 0
