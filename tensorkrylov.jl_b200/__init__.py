"""tensorkrylov.jl_b200 -- B200-native drop-in for TensorKrylov.jl's `tensorkrylov!` solve path.

Layout: csrc/ (CUDA kernels + the C-ABI, built into libtensorkrylov_b200.so),
_capi.py (ctypes binding of include/tensorkrylov_b200.h), api.py (host-side
mirror of the reference's Julia interface), data/ (packed exponential-sum
tables), julia/ (the ccall wrapper a Julia user loads instead), experiments.py (host mirror of the
reference's experiment drivers and the decoder of its stored results).

The directory name contains a dot, so load it with `__graft_entry__.load_package()`
(it registers the module as `tensorkrylov_jl_b200`).
"""
from ._capi import (TK_BREAKDOWN, TK_CONVERGED, TK_FLAG_FIXED_ITERATIONS, TK_FLAG_REFERENCE_H1,  # noqa: F401
                    TK_FLAG_TIME_KERNELS, TK_FLAG_TIME_ALL, TK_NAN, TK_NMAX, TKError, EXPORTS, LIB_PATH, TABLES_PATH,
                    device_count, load_tables)
from . import experiments  # noqa: F401  (host mirror of experiments/*.jl + the stored-result decoder)
from .api import *  # noqa: F401,F403
from .api import (ConvDiff, ConvergenceData, EigValMat, KronMat, KroneckerMatrix, KruskalTensor, Laplace,  # noqa: F401
                  LaplaceDense, NonSymInstance, RandSPD, Solver, SymInstance, TensorArnoldi, TensorLanczos,
                  TensorLanczosReorth, TensorizedSystem, assemble_matrix, kroneckervectorize, nonsym_coefficients,
                  partition_modes, random_rhs, solve_tensorized_system, sym_lookup, tensorkrylov,
                  tridiag_eig_batched)
