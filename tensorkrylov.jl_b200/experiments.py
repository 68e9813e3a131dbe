"""Host-side mirror of the reference's experiment drivers (`experiments/*.jl`, module NumericalExperiments), minus the
plots: the same `Experiment` record, the same operator families and the same entry points, with every solve going
through `solve_tensorized_system` -> libtensorkrylov_b200.so.  Citations are `experiments/<file>:<line>` of
thbake/TensorKrylov.jl.

Also here: the decoder of the reference's stored results (`experiments/data/**`, Julia `Serialization` dumps of
`Experiment` objects written by `serialize_to_file`, experiment_common.jl:115-128), so a stored Julia run can be laid
next to a run of this library without Julia.
"""
from __future__ import annotations

import re
import struct

import numpy as np
import scipy.sparse as sp

from .api import (ConvDiff, ConvergenceData, EigValMat, KroneckerMatrix, Laplace, LaplaceDense, NonSymInstance, RandSPD, SymInstance,
                  TensorArnoldi, TensorizedSystem, TensorLanczos, TensorLanczosReorth, assemble_matrix, random_rhs,
                  solve_tensorized_system)


def multiple_rhs(dims, n, rng=None):
    """system.jl:13: one `random_rhs` per dimension count."""
    return [random_rhs(d, n, rng) for d in dims]


class Experiment:
    """experiment_common.jl:14-42.  `rhs_vec[i]` is the KronProd for `dims[i]`; `conv_vector[i]` its ConvergenceData."""

    def __init__(self, dims, n, nmax, instance, matrixclass, orth_method, rhs_vec):
        self.dims, self.matrixsize, self.nmax = list(dims), n, nmax
        self.instance, self.matrixclass, self.orth_method = instance, matrixclass, orth_method
        self.rhs_vec = rhs_vec
        self.conv_vector = [ConvergenceData(nmax) for _ in self.dims]

    def __len__(self):
        return len(self.dims)

    def __repr__(self):
        return f"Experiment: dimensions d = {self.dims} with matrix size n = {self.matrixsize}"


def _solve(experiment, i, A, tol, **kw):
    system = TensorizedSystem(experiment.instance, A, experiment.rhs_vec[i])
    experiment.conv_vector[i] = solve_tensorized_system(system, experiment.nmax, experiment.orth_method, tol, **kw)


def run_experiments(experiment, tol=1e-9, parameter=None, verbose=False, **kw):
    """`run_experiments!(experiment, tol)` (experiment_common.jl:50-75): the gallery matrix of the experiment's class in
    every mode; with `parameter`, `run_experiments!(experiment, parameter, tol)` (parameterized_systems.jl:32-54): the
    parameterised matrix in every mode, tagged with the experiment's class."""
    A1 = None if parameter is None else parameterize(parameter, experiment.instance)
    for i, d in enumerate(experiment.dims):
        if verbose:
            print(f"d = {d}")
        if A1 is None:
            A = KroneckerMatrix.gallery(experiment.instance, d, experiment.matrixsize, experiment.matrixclass)
        else:
            A = KroneckerMatrix(experiment.instance, [A1] * d, experiment.matrixclass)   # KronMat{U}(A_s, d) + class tag
        _solve(experiment, i, A, tol, verbose=verbose, **kw)
    return experiment


def get_iterations(experiment): return [c.iterations for c in experiment.conv_vector]
def get_max_iteration(experiment): return max(c.niterations for c in experiment.conv_vector)
def get_relative_residuals(experiment): return [c.relative_residual_norm for c in experiment.conv_vector]
def get_projected_residuals(experiment): return [c.projected_residual_norm for c in experiment.conv_vector]
def get_convergence_data(experiment): return experiment.conv_vector


# ---- reproduction.jl ---------------------------------------------------------------------------------------
def reproduce(n=200, tol=1e-9, dims=(5, 10, 50, 100), nmax=None, rhs=None, rng=None, **kw):
    """reproduction.jl:9-21: Laplace / TensorLanczosReorth and ConvDiff / TensorArnoldi on the same right-hand sides,
    nmax = n.  Returns (spd, nonsym)."""
    nmax = n if nmax is None else nmax
    b = multiple_rhs(dims, n, rng) if rhs is None else rhs
    spd = Experiment(dims, n, nmax, SymInstance, Laplace, TensorLanczosReorth, b)
    nonsym = Experiment(dims, n, nmax, NonSymInstance, ConvDiff, TensorArnoldi, b)
    run_experiments(spd, tol, **kw)
    run_experiments(nonsym, tol, **kw)
    return spd, nonsym


# ---- parameterized_systems.jl ------------------------------------------------------------------------------
def parameterize(parameter, instance, n=200):
    """parameterized_systems.jl:3-21.  SymInstance: inv(h^2) SymTridiagonal(alpha, -1); NonSymInstance: the Laplacian plus
    (10/(4h)) diagm(-1 => 1, 0 => 3, 1 => beta, 2 => 1).  Sparse (the driver calls `sparse(...)`, :36)."""
    h = 1.0 / (n + 1)
    inv_h2 = 1.0 / (h * h)
    if instance is SymInstance:
        return (sp.diags([-np.ones(n - 1), parameter * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]) * inv_h2).tocsc()
    lap = sp.diags([-np.ones(n - 1), 2.0 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]) * inv_h2
    conv = sp.diags([np.ones(n - 1), 3.0 * np.ones(n), parameter * np.ones(n - 1), np.ones(n - 2)], [-1, 0, 1, 2])
    return (lap + conv * (10.0 * (1.0 / (4.0 * h)))).tocsc()


def parameterized_experiment(alpha, beta, tol=1e-9, dims=(5, 10, 50, 100), nmax=None, rhs=None, rng=None, **kw):
    """parameterized_systems.jl:56-71: n = 200, nmax = n; the symmetric family runs under the RandSPD class tag (its
    spectral data then comes from `eigvals` of the minors, eigenvalues.jl:337)."""
    n = 200
    nmax = n if nmax is None else nmax
    b = multiple_rhs(dims, n, rng) if rhs is None else rhs
    spd = Experiment(dims, n, nmax, SymInstance, RandSPD, TensorLanczosReorth, b)
    nonsym = Experiment(dims, n, nmax, NonSymInstance, ConvDiff, TensorArnoldi, b)
    run_experiments(spd, tol, parameter=alpha, **kw)
    run_experiments(nonsym, tol, parameter=beta, **kw)
    return spd, nonsym


# ---- eigenvalue_distribution.jl ----------------------------------------------------------------------------
def clusterzero(n):
    """eigenvalue_distribution.jl:110-116: j^2 * inv(n^2) -- the multiplication by the reciprocal is part of the data:
    kappa = k^2 lands on either side of a table row boundary depending on the last bit."""
    kappa = n * n
    return np.array([(j * j) * (1.0 / kappa) for j in range(1, n + 1)], dtype=np.float64)


def clusterone(n):
    """eigenvalue_distribution.jl:118-133"""
    v = np.zeros(n)
    v[0] = 1.0 / (n * n)
    tmp = np.log(float(n))
    for j in range(2, n + 1):
        v[j - 1] = np.log(float(j)) * (1.0 / tmp)
    return v


def perturb_matrix(A, eps):
    """`perturb_matrix!` (eigenvalue_distribution.jl:61-71): A[s] = (s * eps) .+ A[s] -- a broadcast over EVERY entry, so
    mode s becomes a different DENSE matrix."""
    A.M = [np.asarray(M.toarray() if sp.issparse(M) else M, dtype=np.float64) + (s * eps) for s, M in enumerate(A.M, start=1)]


class EigValDist:
    """eigenvalue_distribution.jl:9-33: an Experiment of class EigValMat under TensorLanczosReorth plus its eigenvalues."""

    def __init__(self, dims, eigenvalues, nmax, rhs):
        self.eigenvalues = np.asarray(eigenvalues, dtype=np.float64)
        self.experiment = Experiment(dims, len(self.eigenvalues), nmax, SymInstance, EigValMat, TensorLanczosReorth, rhs)


def run_eigenvalue_experiments(distexp, eps=0.0, tol=1e-9, perturb=False, **kw):
    """`run_experiments!(distexp, eps, tol)` (eigenvalue_distribution.jl:74-107).  The shipped driver calls
    `perturb_matrix!` only when `eps == 0.0` (:92-96), which leaves every matrix as it is, so `perturb=False` (the
    default) IS the shipped behaviour for any eps.  The stored `d2*/d5*/d14*` results were produced by a version that
    perturbed for eps != 0: `perturb=True` reproduces those (tests/golden/corpus.npz)."""
    ex = distexp.experiment
    for i, d in enumerate(ex.dims):
        A = KroneckerMatrix(ex.instance, [assemble_matrix(distexp.eigenvalues, EigValMat)] * d, ex.matrixclass)
        if perturb:
            perturb_matrix(A, eps)
        _solve(ex, i, A, tol, **kw)
    return distexp


def eigenvalue_experiment(n, b, eps=0.0, tol=1e-9, dims=(5, 10, 50, 100), nmax=None, perturb=False, **kw):
    """eigenvalue_distribution.jl:135-153: both clustered spectra on the same right-hand sides, nmax = n."""
    nmax = n if nmax is None else nmax
    distzero = EigValDist(dims, clusterzero(n), nmax, b)
    distone = EigValDist(dims, clusterone(n), nmax, b)
    run_eigenvalue_experiments(distzero, eps, tol, perturb, **kw)
    run_eigenvalue_experiments(distone, eps, tol, perturb, **kw)
    return distzero, distone


def uniform_eigenvalues(n, d, interval):
    """eigenvalue_distribution.jl:173-188: row s = ((s-1) * stepsize / d) .+ LinRange(interval..., n)."""
    ev = np.linspace(interval[0], interval[1], n)
    step = ev[1] - ev[0]
    return np.stack([((s * step) * (1.0 / d)) + ev for s in range(d)])


def uniform_kroneckersum(n, d, interval):
    """eigenvalue_distribution.jl:156-171: a different diagonal matrix in every mode."""
    return KroneckerMatrix(SymInstance, [np.diag(row) for row in uniform_eigenvalues(n, d, interval)], EigValMat)


def uniform_experiment(dims, n, b, interval, tol=1e-9, nmax=None, **kw):
    """eigenvalue_distribution.jl:200-233"""
    dist = EigValDist(dims, np.zeros(n), n if nmax is None else nmax, b)
    ex = dist.experiment
    for i, d in enumerate(ex.dims):
        _solve(ex, i, uniform_kroneckersum(n, d, interval), tol, **kw)
    return dist


# ---- stored results ------------------------------------------------------------------------------------------
_NAMES = [b"NonSymInstance", b"SymInstance", b"LaplaceDense", b"Laplace", b"ConvDiff", b"RandSPD", b"EigValMat",
          b"TensorLanczosReorth", b"TensorLanczos", b"TensorArnoldi"]
_TYPES = {"SymInstance": SymInstance, "NonSymInstance": NonSymInstance, "Laplace": Laplace, "LaplaceDense": LaplaceDense, "ConvDiff": ConvDiff,
          "RandSPD": RandSPD, "EigValMat": EigValMat, "TensorLanczos": TensorLanczos,
          "TensorLanczosReorth": TensorLanczosReorth, "TensorArnoldi": TensorArnoldi}


def julia_arrays(raw):
    """Every 1-d Float64 / Int64 array of a Julia `Serialization` stream, in stream order: (offset, 'f' | 'i', array).
    A 1-d array is `0x15 0x00 <eltype tag> <length> <little-endian data>`, eltype tag 0x0e = Float64, 0x08 = Int64, and
    <length> is 0x31 + int32, 0x06/0x07 + one byte, or a single byte 0xdf + n for n <= 32."""
    out = []
    for m in re.finditer(rb"\x15\x00([\x0e\x08])", raw):
        p, et, tag = m.end(), m.group(1), raw[m.end()]
        if tag == 0x31:
            n = struct.unpack("<i", raw[p + 1:p + 5])[0]
            p += 5
        elif tag in (0x06, 0x07):
            n = raw[p + 1]
            p += 2
        elif 0xDF <= tag <= 0xFF:
            n = tag - 0xDF
            p += 1
        else:
            continue
        if n < 0 or p + 8 * n > len(raw):
            continue
        out.append((m.start(), "f" if et == b"\x0e" else "i",
                    np.frombuffer(raw[p:p + 8 * n], dtype="<f8" if et == b"\x0e" else "<i8").copy()))
    return out


def deserialize_from_file(path, n=200):
    """`deserialize_from_file` (experiment_common.jl:130-143) without Julia: rebuilds the `Experiment` a stored file
    holds -- dims, type tags, the right-hand side of every dimension count and the ConvergenceData histories.  Relies on
    the field order of the struct (dims, ..., rhs_vec, conv_vector) only.  The four oldest files carry three spectral
    vectors (lambda_min, lambda_max, kappa) between the projected residual and the orthogonality history; they are
    skipped."""
    raw = open(path, "rb").read()
    found = set(m.decode() for m in re.findall(b"(" + b"|".join(_NAMES) + b")", raw))
    arrs = julia_arrays(raw)
    if not arrs or arrs[0][1] != "i":
        raise ValueError(f"{path}: not a serialized Experiment")
    dims = [int(x) for x in arrs[0][2]]
    ints = [(o, a) for o, k, a in arrs if k == "i"][1:]
    floats = [(o, a) for o, k, a in arrs if k == "f"]
    rhs_all = [a for o, a in floats if len(a) == n and o < ints[0][0]]
    if len(rhs_all) != sum(dims):
        raise ValueError(f"{path}: expected {sum(dims)} right-hand sides of length {n}, found {len(rhs_all)}")
    inst = "NonSymInstance" if "NonSymInstance" in found else "SymInstance"
    cls = next(c for c in ("LaplaceDense", "ConvDiff", "RandSPD", "EigValMat", "Laplace") if c in found)
    orth = next(c for c in ("TensorArnoldi", "TensorLanczosReorth", "TensorLanczos") if c in found)
    offs = np.concatenate([[0], np.cumsum(dims)])
    rhs_vec, convs = [], []
    for i, d in enumerate(dims):
        rhs_vec.append(rhs_all[offs[i]:offs[i + 1]])
        o_it, iters = ints[i]
        o_next = ints[i + 1][0] if i + 1 < len(ints) else len(raw)
        hist = [a for o, a in floats if o_it < o < o_next and len(a) == len(iters)]
        if len(hist) not in (3, 6):
            raise ValueError(f"{path}: unexpected ConvergenceData layout for d = {d}")
        cd = ConvergenceData(len(iters))
        cd.iterations = iters
        cd.relative_residual_norm, cd.projected_residual_norm, cd.orthogonality_data = hist[0], hist[1], hist[-1]
        convs.append(cd)
    nmax = max(len(c.iterations) for c in convs)
    ex = Experiment(dims, n, nmax, _TYPES[inst], _TYPES[cls], _TYPES[orth], rhs_vec)
    ex.conv_vector = convs
    return ex


def serialize_to_file(path, experiment):
    """`serialize_to_file` (experiment_common.jl:115-128) as a compressed .npz (numpy has no use for Julia's format)."""
    out = {"dims": np.array(experiment.dims), "matrixsize": experiment.matrixsize, "nmax": experiment.nmax,
           "tags": np.array([experiment.instance.__name__, experiment.matrixclass.__name__, experiment.orth_method.__name__])}
    for d, rhs, c in zip(experiment.dims, experiment.rhs_vec, experiment.conv_vector):
        out[f"rhs_d{d}"] = np.asarray(rhs[0])
        out[f"niterations_d{d}"] = c.niterations
        out[f"iterations_d{d}"] = c.iterations
        out[f"relres_d{d}"] = c.relative_residual_norm
        out[f"projres_d{d}"] = c.projected_residual_norm
        out[f"orth_d{d}"] = c.orthogonality_data
    np.savez_compressed(path, **out)
