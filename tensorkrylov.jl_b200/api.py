"""Host-side mirror of the reference's Julia interface for the solve path.

Same names, argument meaning and exits as the reference (citations are
`src/<file>:<line>` of thbake/TensorKrylov.jl); all arithmetic of the path runs
in libtensorkrylov_b200.so on the GPU.  What stays on the host is exactly what
stays in Julia for the `ccall` wrapper (julia/TensorKrylovB200.jl): building
the synthetic operators, normalising b, and -- for matrix classes without an
analytic spectrum -- the eigen-extremes of the minors of A_1, which depend only
on A and are fed to the library as the exp-sum schedule.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _capi
from ._capi import TKError, check, dptr, lib


# ---- type tags (tensor_struct.jl:18-23, 83-85; decompositions.jl:120-176) -------------------
class Instance: code = None
class SymInstance(Instance): code = _capi.TK_SYM
class NonSymInstance(Instance): code = _capi.TK_NONSYM

class MatrixGallery: code = _capi.TK_GENERIC
class LaplaceDense(MatrixGallery): code = _capi.TK_LAPLACE_DENSE
class Laplace(MatrixGallery): code = _capi.TK_LAPLACE
class ConvDiff(MatrixGallery): code = _capi.TK_CONVDIFF
class EigValMat(MatrixGallery): code = _capi.TK_EIGVALMAT
class RandSPD(MatrixGallery): code = _capi.TK_RANDSPD

class TensorDecomposition: code = None
class TensorLanczos(TensorDecomposition): code = _capi.TK_LANCZOS
class TensorLanczosReorth(TensorDecomposition): code = _capi.TK_LANCZOS_REORTH
class TensorArnoldi(TensorDecomposition): code = _capi.TK_ARNOLDI


class CompressedNormBreakdown(ArithmeticError):
    """utils.jl:7-14"""
    def __init__(self, r_comp):
        super().__init__(f"compressed residual norm {r_comp} < 0")
        self.r_comp = r_comp


# ---- synthetic operators (tensor_struct.jl:48-79) -------------------------------------------
def assemble_matrix(n, cls, c=10.0, rng=None):
    """Laplace/ConvDiff -> scipy CSC (Julia's SparseMatrixCSC); LaplaceDense/RandSPD/EigValMat -> dense."""
    if cls is EigValMat:
        return np.diag(np.asarray(n, dtype=np.float64))                     # n is the eigenvalue vector (:70)
    h = 1.0 / (n + 1)
    inv_h2 = 1.0 / (h * h)
    lap = sp.diags([-np.ones(n - 1), 2.0 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]) * inv_h2
    if cls is Laplace:
        return lap.tocsc()
    if cls is LaplaceDense:
        return lap.toarray()
    if cls is ConvDiff:
        f = c * (1.0 / (4.0 * h))
        conv = sp.diags([np.ones(n - 1), 3.0 * np.ones(n), -5.0 * np.ones(n - 1), np.ones(n - 2)], [-1, 0, 1, 2]) * f
        return (lap + conv).tocsc()
    if cls is RandSPD:
        R = (rng or np.random.default_rng()).random((n, n))
        S = R.T @ R
        return np.tril(S) + np.tril(S, -1).T
    raise TypeError(f"no assemble_matrix method for {cls}")


class KroneckerMatrix:
    """tensor_struct.jl:168-231.  M is the list of d coefficient matrices (aliases allowed)."""

    def __init__(self, instance, M, matrixclass=MatrixGallery):
        self.instance, self.M, self.matrixclass = instance, list(M), matrixclass

    @classmethod
    def gallery(cls, instance, d, n, matrixclass, **kw):
        A = assemble_matrix(n, matrixclass, **kw)      # one object aliased d times (:208-210)
        return cls(instance, [A] * d, matrixclass)

    def __len__(self): return len(self.M)
    def __getitem__(self, s): return self.M[s]
    def dimensions(self): return [A.shape[0] for A in self.M]

KronMat = KroneckerMatrix


class KruskalTensor:
    """tensor_struct.jl:283-316: lambda (t) and factor matrices fmat[s] (n_s x t)."""
    def __init__(self, lambda_, fmat):
        self.lambda_ = np.asarray(lambda_, dtype=np.float64)
        self.fmat = list(fmat)
    def ncomponents(self): return len(self.lambda_)
    def ndims(self): return len(self.fmat)


def kroneckervectorize(x):
    """tensor_struct.jl:361-384 (mode 1 fastest)."""
    out = 0.0
    for i in range(x.ncomponents()):
        tmp = x.fmat[-1][:, i]
        for j in range(x.ndims() - 2, -1, -1):
            tmp = np.kron(tmp, x.fmat[j][:, i])
        out = out + x.lambda_[i] * tmp
    return out


def random_rhs(d, n, rng=None):
    """system.jl:5-11: one U(0,1) vector shared by all d modes."""
    bs = (rng or np.random.default_rng()).random(n)
    return [bs for _ in range(d)]


class TensorizedSystem:
    """system.jl:15-43: asserts the shapes and normalises every b_s."""
    def __init__(self, instance, A, b, normalize_rhs=True):
        assert len(A) == len(b)
        assert all(n == len(bs) for n, bs in zip(A.dimensions(), b))
        self.instance, self.d, self.n, self.A = instance, len(A), A[0].shape[0], A
        if normalize_rhs:
            cache = {}
            out = []
            for bs in b:                         # rhs[i] *= inv(norm(rhs[i])), utils.jl:446-454
                key = id(bs)
                if key not in cache:
                    cache[key] = np.asarray(bs, dtype=np.float64) * (1.0 / np.linalg.norm(bs))
                out.append(cache[key])
            b = out
        self.b = [np.asarray(bs, dtype=np.float64) for bs in b]


class ConvergenceData:
    """convergence.jl:3-32"""
    def __init__(self, nmax):
        self.niterations = nmax
        self.iterations = np.arange(1, nmax + 1)
        self.relative_residual_norm = np.ones(nmax)
        self.projected_residual_norm = np.ones(nmax)
        self.orthogonality_data = np.ones(nmax)
        self.status = None
        self.term_k = None

    def resize(self, k):
        self.iterations = self.iterations[:k]
        self.relative_residual_norm = self.relative_residual_norm[:k]
        self.projected_residual_norm = self.projected_residual_norm[:k]
        self.orthogonality_data = self.orthogonality_data[:k]


# ---- spectral data / schedule ----------------------------------------------------------------
def extreme_eigvals(A1, d, k, instance, matrixclass):
    """eigenvalues.jl:335-350 for the classes without an analytic formula (host LAPACK, depends only on A_1)."""
    if instance is NonSymInstance:
        M = A1[:k, :k].toarray() if sp.issparse(A1) else np.asarray(A1)[:k, :k]
        ev = np.linalg.eigvals(M)
        if np.abs(ev.imag).max() > 0:
            raise ArithmeticError("complex eigenvalues in a minor of A_1")
        return float(ev.real.min()) * d, None
    if matrixclass is Laplace:
        lmin, lmax = C.c_double(), C.c_double()
        check(lib.tk_laplace_extremes(d, A1.shape[0], k, C.byref(lmin), C.byref(lmax)))
        return lmin.value, lmax.value
    if matrixclass is RandSPD:
        M = A1[:k, :k].toarray() if sp.issparse(A1) else np.asarray(A1)[:k, :k]
        ev = np.linalg.eigvalsh(M)
        return float(ev.min()) * d, float(ev.max()) * d
    if matrixclass is EigValMat:
        dg = (A1.diagonal() if sp.issparse(A1) else np.diag(np.asarray(A1)))[:k]
        return float(dg.min()) * d, float(dg.max()) * d
    raise TypeError(f"no extreme_eigvals method for ({instance.__name__}, {matrixclass.__name__})")


def sym_lookup(kappa, tol):
    """ApproximationData lookup (approximation.jl:65-84, 119-147) through the library's tables."""
    _capi.load_tables()
    t, dg, od = C.c_int32(), C.c_int32(), C.c_int32()
    om, al = np.zeros(64), np.zeros(64)
    check(lib.tk_tables_sym_lookup(kappa, tol, C.byref(t), C.byref(dg), C.byref(od), dptr(om), dptr(al)))
    return t.value, om[:t.value].copy(), al[:t.value].copy(), dg.value, od.value


def sym_rank_coefficients(kappa, rank):
    """The coefficient file of a given rank in kappa's table row (approximation.jl:119-147): (omega, alpha, error)."""
    _capi.load_tables()
    om, al, err = np.zeros(64), np.zeros(64), C.c_double()
    check(lib.tk_tables_sym_rank(kappa, rank, dptr(om), dptr(al), C.byref(err)))
    return om[:rank].copy(), al[:rank].copy(), err.value


def nonsym_coefficients(lambda_min, tol):
    """approximation.jl:86-107, 150-158"""
    cap = 8192
    rank, nt = C.c_int32(), C.c_int32()
    om, al = np.zeros(cap), np.zeros(cap)
    check(lib.tk_nonsym_coefficients(lambda_min, tol, cap, C.byref(rank), C.byref(nt), dptr(om), dptr(al)))
    return rank.value, om[:nt.value].copy(), al[:nt.value].copy()


# where Solver.set_schedule takes the eigen-extremes of the minors of A_1 from: "library" (tk_schedule) or "lapack"
DEFAULT_SPECTRAL = "library"


# ---- the device-resident solver --------------------------------------------------------------
class Solver:
    """One tk_handle: the state tensorkrylov! keeps in `tensor_decomp`, b~, spectraldata, approxdata."""

    def __init__(self, d, n, nmax, instance, matrixclass, variant, flags=_capi.TK_FLAG_REFERENCE_H1,
                 device=0, rank=0, world=1, unique_id=None):
        self.d, self.n, self.nmax = d, n, nmax
        self.instance, self.matrixclass, self.variant = instance, matrixclass, variant
        ns = (C.c_int64 * d)(*([n] * d))
        h = C.c_void_p()
        uid = C.c_char_p(unique_id) if unique_id is not None else None
        check(lib.tk_create(C.byref(h), d, ns, nmax, instance.code, matrixclass.code, variant.code, flags,
                            device, rank, world, uid))
        self.h = h
        f, c = C.c_int32(), C.c_int32()
        check(lib.tk_local_modes(self.h, C.byref(f), C.byref(c)))
        self.first, self.count = f.value, c.value
        # global modes whose inputs this rank wants (its own, plus mode 0 under TK_FLAG_REFERENCE_H1)
        self.fed = list(range(self.first, self.first + self.count))
        need0 = C.c_int32()
        check(lib.tk_needs_mode(self.h, 0, C.byref(need0)))
        if need0.value and 0 not in self.fed:
            self.fed.insert(0, 0)

    def close(self):
        if getattr(self, "h", None):
            lib.tk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # inputs
    def set_operators(self, M):
        seen = {}
        one = bool(self.fed) and all(M[s] is M[self.fed[0]] for s in self.fed)
        for s in self.fed:
            A = M[s]
            if one and seen:          # one matrix object in every slot (tensor_struct.jl:208-210)
                check(lib.tk_share_operator_all(self.h, self.fed[0]))
                return
            if id(A) in seen:
                check(lib.tk_share_operator(self.h, s, seen[id(A)]))
                continue
            self._feed_operator(s, A)
            seen[id(A)] = s

    def set_rhs(self, b):
        first = b[self.fed[0]] if self.fed else None
        if self.fed and all(b[s] is first for s in self.fed):
            v = _capi.as_f64(first)
            check(lib.tk_set_rhs_all(self.h, dptr(v), len(v)))
            return
        for s in self.fed:
            v = _capi.as_f64(b[s])
            check(lib.tk_set_rhs(self.h, s, dptr(v), len(v)))

    def _feed_operator(self, s, A):
        if sp.issparse(A):
            Ac = A.tocsc()
            Ac.sort_indices()
            colptr = np.ascontiguousarray(Ac.indptr, dtype=np.int64) + 1      # Julia is 1-based
            rowval = np.ascontiguousarray(Ac.indices, dtype=np.int64) + 1
            nz = np.ascontiguousarray(Ac.data, dtype=np.float64)
            check(lib.tk_set_operator_csc(self.h, s, Ac.shape[0],
                                          colptr.ctypes.data_as(C.POINTER(C.c_int64)),
                                          rowval.ctypes.data_as(C.POINTER(C.c_int64)), dptr(nz)))
        else:
            Af = np.asfortranarray(A, dtype=np.float64)
            check(lib.tk_set_operator_dense(self.h, s, Af.shape[0], dptr(Af), b"F"))

    def set_schedule(self, A1, tol, spectral=None):
        """The two update_data! calls of every iteration (tensor_krylov_method.jl:72-73), hoisted out of the loop.

        spectral="library" (default): tk_schedule -- the eigen-extremes of the leading minors of A_1
        (eigenvalues.jl:335-350) are computed inside the library on the host threads.  spectral="lapack": they come
        from numpy's LAPACK here (`extreme_eigvals`) and are fed through tk_set_schedule, the way the Julia wrapper
        feeds Julia's own SpectralData; the two agree to eps * cond(minor) in lambda_min."""
        _capi.load_tables()
        spectral = spectral or DEFAULT_SPECTRAL
        if self.instance is SymInstance and self.matrixclass is Laplace:
            check(lib.tk_schedule_laplace(self.h, tol))
            return
        if spectral == "library":
            if 0 not in self.fed:
                self._feed_operator(0, A1)       # only its leading block is kept on a rank that does not own mode 1
            check(lib.tk_schedule(self.h, tol))
            return
        for k in range(2, self.nmax + 1):
            lmin, lmax = extreme_eigvals(A1, self.d, k, self.instance, self.matrixclass)
            if self.instance is SymInstance:
                kappa = lmax * (1.0 / lmin)
                t, om, al, _, _ = sym_lookup(kappa, tol)
            else:
                _, om, al = nonsym_coefficients(lmin, tol)
                t = len(om)
            check(lib.tk_set_schedule(self.h, k, lmin, t, dptr(al), dptr(om)))

    def set_schedule_entry(self, k, lambda_min, alpha, omega):
        al, om = _capi.as_f64(alpha), _capi.as_f64(omega)
        check(lib.tk_set_schedule(self.h, k, lambda_min, len(al), dptr(al), dptr(om)))

    # the solve
    def solve(self, tol):
        nmax = self.nmax
        st, nit, tk = C.c_int32(), C.c_int64(), C.c_int32()
        rr, pr, ol = np.empty(nmax), np.empty(nmax), np.empty(nmax)
        check(lib.tk_solve(self.h, tol, C.byref(st), C.byref(nit), C.byref(tk), dptr(rr), dptr(pr), dptr(ol)))
        return dict(status=st.value, niterations=nit.value, term_k=tk.value, relres=rr, projres=pr, orth=ol)

    def solution_rank(self):
        t = C.c_int32()
        check(lib.tk_solution_rank(self.h, C.byref(t)))
        return t.value

    def solution(self, force=False, pinned=False):
        """x of the iteration the loop left at (basis_tensor_mul!, utils.jl:478-488): (lambda, {mode: n x t matrix}).
        pinned=True puts the factor matrices into page-locked memory (DMA without a staging copy); the returned
        matrices are then views of one PinnedArray kept alive in `self.pinned_result`."""
        t = self.solution_rank()
        lam = np.zeros(max(t, 1))
        shape = (max(self.count, 1), t, self.n)      # [mode][column][row] = n x t column-major per mode
        if pinned:
            self.pinned_result = _capi.PinnedArray(shape)
            buf = self.pinned_result.array
        else:
            buf = np.empty(shape)
        check(lib.tk_get_solution_all(self.h, dptr(lam), len(lam), dptr(buf), buf.size, 1 if force else 0))
        fmat = {self.first + i: buf[i].T for i in range(self.count)}
        return lam[:t], fmat

    def solution_mode(self, s, force=False):
        t = self.solution_rank()
        lam = np.zeros(max(t, 1))
        F = np.empty((t, self.n))
        check(lib.tk_get_solution(self.h, s, dptr(lam), len(lam), dptr(F), F.size, 1 if force else 0))
        return lam[:t], F.T

    def solution_device(self, dev_ptr, capacity, force=False):
        """Factor matrices into caller-owned DEVICE memory ([mode][t][n] doubles at dev_ptr); returns lambda."""
        t = self.solution_rank()
        lam = np.zeros(max(t, 1))
        check(lib.tk_get_solution_device(self.h, dptr(lam), len(lam), C.c_void_p(dev_ptr), capacity, 1 if force else 0))
        return lam[:t]

    def detail(self, k0=2, k1=None):
        """Per-iteration terms of the residual estimate of the last solve, rows k0..k1."""
        k1 = self.nmax if k1 is None else k1
        out = np.zeros((k1 - k0 + 1, 8))
        check(lib.tk_get_detail(self.h, k0, k1, dptr(out)))
        names = ("hy2", "hyb", "bb", "boundary", "r_comp", "r_norm", "t", "lambda_min")
        return {nm: out[:, i].copy() for i, nm in enumerate(names)}

    def solve_info(self):
        g, ms, px, sg = C.c_int32(), C.c_double(), C.c_int32(), C.c_int32()
        check(lib.tk_get_solve_info(self.h, C.byref(g), C.byref(ms), C.byref(px), C.byref(sg)))
        return dict(graphs_launched=g.value, graph_build_ms=ms.value, peer_exchange=bool(px.value), segments=sg.value)

    # test-only phases and readers
    def begin(self): check(lib.tk_begin(self.h))
    def step_bases(self, k): check(lib.tk_step_bases(self.h, k))
    def compress(self, k): check(lib.tk_compress(self.h, k))

    def residual(self, k, tol):
        out = np.zeros(8)
        check(lib.tk_residual(self.h, k, tol, dptr(out)))
        return dict(hy2=out[0], hyb=out[1], bb=out[2], boundary=out[3], r_comp=out[4], r_norm=out[5],
                    t=int(out[6]), lambda_min=out[7])

    def get_H(self, s):
        nc = self.nmax + 1
        H = np.zeros((nc, nc), order="F")
        check(lib.tk_get_H(self.h, s, dptr(H)))
        return H

    def get_V(self, s, col):
        v = np.zeros(self.n)
        check(lib.tk_get_V(self.h, s, col, dptr(v)))
        return v

    def get_bt(self, s):
        bt = np.zeros(self.nmax + 1)
        check(lib.tk_get_bt(self.h, s, dptr(bt)))
        return bt

    def get_Y(self, s, k):
        t = C.c_int32()
        check(lib.tk_get_Y(self.h, s, k, None, C.byref(t)))
        Y = np.zeros((k, t.value), order="F")
        check(lib.tk_get_Y(self.h, s, k, dptr(Y), C.byref(t)))
        return Y

    def get_eig(self, s, k):
        th = np.zeros(k)
        Q = np.zeros((k, k), order="F")
        check(lib.tk_get_eig(self.h, s, k, dptr(th), dptr(Q)))
        return th, Q

    def orth_state(self, s):
        S, fb = C.c_double(), C.c_int32()
        check(lib.tk_get_orth_state(self.h, s, C.byref(S), C.byref(fb)))
        return S.value, fb.value

    def timing_mark(self):
        check(lib.tk_timing_mark(self.h))

    def timing(self, which):
        ms, n, by = C.c_double(), C.c_int64(), C.c_double()
        check(lib.tk_get_timing(self.h, which, C.byref(ms), C.byref(n), C.byref(by)))
        return ms.value, n.value, by.value

    def launch_count(self):
        n = C.c_int64()
        check(lib.tk_launch_count(self.h, C.byref(n)))
        return n.value


def tridiag_eig_batched(diag, sub, device=0, vectors=True, info=None):
    """Kernel (2) on its own: diag (nb, k), sub (nb, k-1) -> theta (nb, k), Q (nb, k, k) with Q[p][:, i] eigenvector i.
    info (a dict) receives the number of problems the QL fallback had to redo."""
    diag = _capi.as_f64(np.atleast_2d(diag))
    nb, k = diag.shape
    sub = _capi.as_f64(np.atleast_2d(sub)) if k > 1 else np.zeros((nb, 0))
    theta = np.zeros((nb, k))
    Q = np.zeros((nb, k, k)) if vectors else None
    fb = C.c_int32(0)
    check(lib.tk_tridiag_eig_batched(device, nb, k, dptr(diag), dptr(sub) if k > 1 else None, dptr(theta),
                                     dptr(Q) if vectors else None, C.byref(fb)))
    if info is not None:
        info["fallbacks"] = fb.value
    if vectors:
        Q = np.transpose(Q, (0, 2, 1)).copy()   # stored column-major per problem
    return theta, Q


# ---- the reference's entry points ------------------------------------------------------------
def tensorkrylov(convergence_data, A, b, tol, nmax, orthonormalization_type, flags=_capi.TK_FLAG_REFERENCE_H1,
                 device=0, verbose=True, solver_out=None):
    """`tensorkrylov!` (tensor_krylov_method.jl:36-125).  Returns the KruskalTensor x on convergence, else None;
    fills `convergence_data` exactly like the reference (ones, then entries 2..k; resize! on breakdown)."""
    d, n = len(A), A[0].shape[0]
    slv = Solver(d, n, nmax, A.instance, A.matrixclass, orthonormalization_type, flags=flags, device=device)
    try:
        slv.set_operators(A.M)
        slv.set_rhs(b)
        slv.set_schedule(A[0], tol)
        res = slv.solve(tol)
        cd = convergence_data
        cd.status, cd.term_k = res["status"], res["term_k"]
        cd.relative_residual_norm[:] = res["relres"]
        cd.projected_residual_norm[:] = res["projres"]
        cd.orthogonality_data[:] = res["orth"]
        if res["status"] == _capi.TK_BREAKDOWN:
            if verbose:
                print(f"Early termination at k = {res['term_k']} due to compressed norm breakdown")
            cd.niterations = res["niterations"]
            cd.resize(cd.niterations)
            return None
        if res["status"] == _capi.TK_CONVERGED:
            lam, fmat = slv.solution()
            if verbose:
                print("Convergence")
            return KruskalTensor(lam, [fmat[s] for s in range(d)])
        if res["status"] == _capi.TK_NAN:
            raise ArithmeticError(f"NaN in the residual estimate at k = {res['term_k']}")
        if verbose:
            print("No convergence")
        return None
    finally:
        if solver_out is not None:
            solver_out.append(slv)
        else:
            slv.close()


def solve_tensorized_system(system, nmax, orthogonalization_type, tol=1e-9, **kw):
    """system.jl:65-83"""
    convergencedata = ConvergenceData(nmax)
    tensorkrylov(convergencedata, system.A, system.b, tol, nmax, orthogonalization_type, **kw)
    return convergencedata


def partition_modes(d, world, chunk=16):
    """The block partition tk_create uses: (first, count) per rank, aligned to combine chunks when d allows."""
    per = (d + world - 1) // world
    if d >= world * chunk:
        per = ((per + chunk - 1) // chunk) * chunk
    out = []
    for r in range(world):
        first = min(d, r * per)
        out.append((first, max(0, min(per, d - first))))
    return out
