"""CPU oracle for the tensorized Krylov solve -- TEST INFRASTRUCTURE ONLY.

A numpy/scipy Float64 restatement of the reference algorithm
(thbake/TensorKrylov.jl, `tensorkrylov!`), written from the reference's
behaviour; every function cites the reference file:line it follows (paths are
relative to the reference root).  The reference is pure Julia and Julia is not
available in this image, so the oracle is a port, not the reference itself.

PARITY PIN: this oracle is pinned against the reference's own stored results
(tests/golden/*.npz, decoded from experiments/data/** by tools/make_golden.py)
and the reference's known-answer tests (test/utils.jl:188-227,
test/eigenvalues.jl:5-73); see tests/test_oracle_golden.py.

Third-party arithmetic the reference delegates to Julia 1.9.3 stdlib
(Manifest.toml:3) and that is restated here by its published algorithm:
  * exp(::Symmetric)  = eigendecomposition  V exp(L) V'   -> numpy.linalg.eigh
  * exp(::Matrix)     = Higham Pade-13 scaling & squaring -> scipy.linalg.expm
  * eigvals           = LAPACK                             -> numpy.linalg.eigvals(h)
  * dot/nrm2/syrk/gemm = OpenBLAS 0.3.21                   -> numpy (OpenBLAS)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm
may import this module, and only as the checker / the CPU baseline.  The
product path (tensorkrylov.jl_b200) never imports it.
"""
from __future__ import annotations

import math
import os
import struct

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

# enum values shared with include/tensorkrylov_b200.h
SYM, NONSYM = 0, 1
LAPLACE_DENSE, LAPLACE, CONVDIFF, EIGVALMAT, RANDSPD = 0, 1, 2, 3, 4
LANCZOS, LANCZOS_REORTH, ARNOLDI = 0, 1, 2
ST_CONVERGED, ST_NMAX, ST_BREAKDOWN, ST_NAN = 0, 1, 2, 3

SQRT_EPS = math.sqrt(np.finfo(np.float64).eps)  # orthogonal_bases.jl:123


# ----------------------------------------------------------------------------
# synthetic operators (tensor_struct.jl:48-79)
# ----------------------------------------------------------------------------
def assemble_matrix(n, cls, c=10.0, eigenvalues=None, rng=None):
    """tensor_struct.jl:48-79.  Returns scipy CSR (Laplace, ConvDiff) or a dense array."""
    h = 1.0 / (n + 1)
    inv_h2 = 1.0 / (h * h)
    if cls in (LAPLACE, LAPLACE_DENSE):
        # inv(h^2) * SymTridiagonal(2 ones(n), -ones(n))   (:50-51)
        L = sp.diags([-np.ones(n - 1), 2.0 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1], format="csr") * inv_h2
        return L.toarray() if cls == LAPLACE_DENSE else L.tocsr()
    if cls == CONVDIFF:
        # L + (c/(4h)) * diagm(-1=>1, 0=>3, 1=>-5, 2=>1)   (:60-68)
        L = sp.diags([-np.ones(n - 1), 2.0 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1], format="csr") * inv_h2
        f = c * (1.0 / (4.0 * h))
        C = sp.diags([np.ones(n - 1), 3.0 * np.ones(n), -5.0 * np.ones(n - 1), np.ones(n - 2)],
                     [-1, 0, 1, 2], format="csr") * f
        return (L + C).tocsr()
    if cls == EIGVALMAT:
        return np.diag(np.asarray(eigenvalues, dtype=np.float64))  # :70
    if cls == RANDSPD:
        R = (rng or np.random.default_rng(0)).random((n, n))       # :73-79
        S = R.T @ R
        return np.tril(S) + np.tril(S, -1).T
    raise ValueError("unknown matrix class")


def random_rhs(d, n, seed=12345):
    """system.jl:5-11: ONE U(0,1) vector replicated over the d modes."""
    b = np.random.default_rng(seed).random(n)
    return [b.copy() for _ in range(d)]


def normalize_rhs(bs):
    """utils.jl:446-454 / system.jl:33-37:  b_s <- b_s * inv(norm(b_s))."""
    return [b * (1.0 / np.linalg.norm(b)) for b in bs]


# ----------------------------------------------------------------------------
# exponential-sum tables (approximation.jl:44-158)
# ----------------------------------------------------------------------------
class ExpSumTables:
    """The CSV error table + the (t, digit, order) -> (omega, alpha) coefficient files."""

    def __init__(self, R, err, ranks, coeffs):
        self.R = np.asarray(R, dtype=np.float64)
        self.err = np.asarray(err, dtype=np.float64)
        self.ranks = np.asarray(ranks, dtype=np.int64)
        self.coeffs = coeffs  # {(t, digit, order): (omega, alpha)}

    @classmethod
    def from_packed(cls, path):
        """Format written by tools/pack_tables.py."""
        raw = open(path, "rb").read()
        assert raw[:8] == b"TKXSUM01", "bad table file"
        off = 8
        nrows, nranks = struct.unpack_from("<ii", raw, off); off += 8
        R = np.frombuffer(raw, "<f8", nrows, off); off += 8 * nrows
        err = np.frombuffer(raw, "<f8", nrows * nranks, off).reshape(nrows, nranks); off += 8 * nrows * nranks
        ranks = np.frombuffer(raw, "<i4", nranks, off); off += 4 * nranks
        nfiles, _ = struct.unpack_from("<ii", raw, off); off += 8
        coeffs = {}
        for _ in range(nfiles):
            t, digit, order, _ = struct.unpack_from("<iiii", raw, off); off += 16
            om = np.frombuffer(raw, "<f8", t, off).copy(); off += 8 * t
            al = np.frombuffer(raw, "<f8", t, off).copy(); off += 8 * t
            coeffs[(t, digit, order)] = (om, al)
        return cls(R, err, ranks, coeffs)

    @classmethod
    def from_reference_dir(cls, cdir):
        """Reads coefficients_data/ as shipped (approximation.jl:44-54, 119-147)."""
        import csv
        import re
        with open(os.path.join(cdir, "output_data", "tabelle_complete.csv")) as f:
            rows = [r for r in csv.reader(f) if r]
        ranks = [int(c) for c in rows[0][1:]]
        R = [float(r[0]) for r in rows[1:]]
        err = [[float(x) for x in r[1:]] for r in rows[1:]]
        coeffs = {}
        pat = re.compile(r"^1_xk(\d\d)\.(\d+)_(\d+)$")
        for name in os.listdir(cdir):
            m = pat.match(name)
            if not m:
                continue
            t = int(m.group(1))
            vals = [float(l.split("{")[0]) for l in open(os.path.join(cdir, name)) if l.strip()]
            coeffs[(t, int(m.group(2)), int(m.group(3)))] = (np.array(vals[:t]), np.array(vals[t:2 * t]))
        return cls(R, err, ranks, coeffs)

    @staticmethod
    def parse_condition(kappa):
        """approximation.jl:109-116: kappa rounded DOWN to one significant digit."""
        order = int(math.floor(math.log10(kappa)))
        digit = int(math.floor(kappa / (10.0 ** float(order))))
        return order, digit

    def sym_lookup(self, kappa, tol):
        """approximation.jl:65-84 + 119-147.  Returns (t, omega, alpha, digit, order)."""
        order, digit = self.parse_condition(kappa)
        while True:
            hit = np.nonzero(self.R == digit * 10.0 ** order)[0]   # getclosestrow :56-63
            if len(hit):
                break
            digit += 1                                             # :71-76
            if digit > 1000:
                raise RuntimeError("condition number outside the table")
        row = self.err[hit[0]]
        mask = tol >= row                                          # :79
        if not mask.any():
            raise RuntimeError("no rank reaches the tolerance")
        t = int(self.ranks[mask].min())                            # :80-82
        om, al = self.coeffs[(t, digit, order)]
        return t, om, al, digit, order


def nonsym_coeffs(lambda_min, tol):
    """approximation.jl:86-107 (rank) and 150-158 (sinc-quadrature coefficients)."""
    rank = 1
    bound = lambda r: 2.75 * (1.0 / lambda_min) * math.exp(-math.pi * math.sqrt(r / 2))
    while bound(rank) > tol:
        rank += 1
    h = math.pi * (1.0 / math.sqrt(rank))
    js = range(-rank, rank + 1)
    alpha = np.array([math.log(math.exp(j * h) + math.sqrt(1.0 + math.exp(2 * j * h))) for j in js])
    omega = np.array([h * (1.0 / math.sqrt(1.0 + math.exp(-2 * j * h))) for j in js])
    return rank, omega, alpha


# ----------------------------------------------------------------------------
# spectral data (eigenvalues.jl:247-370): depends only on (A_1, d, k)
# ----------------------------------------------------------------------------
def laplace_extremes(d, n, k):
    """eigenvalues.jl:247-265: eigen-extremes of the k x k leading minor of A_1, times d."""
    h = 1.0 / (n + 1)
    lmin = 4.0 * (1.0 / (h * h)) * math.sin(1 * math.pi * (1.0 / (2 * (k + 1)))) ** 2 * d
    lmax = 4.0 * (1.0 / (h * h)) * math.sin(k * math.pi * (1.0 / (2 * (k + 1)))) ** 2 * d
    return lmin, lmax


def extreme_eigvals(A1, d, k, instance, cls):
    """eigenvalues.jl:335-350.  Returns (lambda_min, lambda_max or None)."""
    n = A1.shape[0]
    if instance == NONSYM:                                          # :344-350
        M = A1[:k, :k].toarray() if sp.issparse(A1) else np.asarray(A1)[:k, :k]
        ev = np.linalg.eigvals(M)
        if np.abs(ev.imag).max() > 0:
            raise ArithmeticError("complex eigenvalues in a minor of A_1 (the reference's minimum() throws)")
        return float(ev.real.min()) * d, None
    if cls == LAPLACE:                                              # :335
        return laplace_extremes(d, n, k)
    if cls == RANDSPD:                                              # :337
        M = A1[:k, :k].toarray() if sp.issparse(A1) else np.asarray(A1)[:k, :k]
        ev = np.linalg.eigvalsh(M)
        return float(ev.min()) * d, float(ev.max()) * d
    if cls == EIGVALMAT:                                            # :339
        dg = (A1.diagonal() if sp.issparse(A1) else np.diag(np.asarray(A1)))[:k]
        return float(dg.min()) * d, float(dg.max()) * d
    raise TypeError("no extreme_eigvals method for this (instance, class) -- the reference throws MethodError")


def build_schedule(A1, d, nmax, tol, instance, cls, tables):
    """For k = 2..nmax: (lambda_min, t, alpha, omega).  Restates the two update_data!
    calls (tensor_krylov_method.jl:72-73), which never look at the Krylov state."""
    sched = {}
    for k in range(2, nmax + 1):
        lmin, lmax = extreme_eigvals(A1, d, k, instance, cls)
        if instance == SYM:
            kappa = lmax * (1.0 / lmin)                             # eigenvalues.jl:360
            t, om, al, digit, order = tables.sym_lookup(kappa, tol)
            sched[k] = dict(lambda_min=lmin, t=t, alpha=al, omega=om, rank=t, kappa=kappa, R=(digit, order))
        else:
            rank, om, al = nonsym_coeffs(lmin, tol)
            sched[k] = dict(lambda_min=lmin, t=len(om), alpha=al, omega=om, rank=rank)
    return sched


# ----------------------------------------------------------------------------
# Krylov steps (orthogonal_bases.jl:15-139); V is n x (nmax+1), H is (nmax+1)^2,
# 1-based k as in the reference.
# ----------------------------------------------------------------------------
def mgs_step(A, V, H, k):
    """orthogonal_bases.jl:15-37 (two Gram-Schmidt passes; no zero-norm guard)."""
    v = A @ V[:, k - 1]
    for i in range(k):
        H[i, k - 1] = v @ V[:, i]
        v = v - H[i, k - 1] * V[:, i]
    for i in range(k):
        H[i, k - 1] += v @ V[:, i]
        v = v - (v @ V[:, i]) * V[:, i]
    H[k, k - 1] = np.linalg.norm(v)
    V[:, k] = v * (1.0 / H[k, k - 1])


def orthogonality_loss(V, k):
    """orthogonal_bases.jl:250-257: || V[:,1:k]' V[:,1:k] - I ||_F."""
    G = V[:, :k].T @ V[:, :k]
    return float(np.linalg.norm(G - np.eye(k)))


def lanczos_step(A, V, H, k, reorth, stats=None):
    """orthogonal_bases.jl:39-67 (TTR) and 98-139 (monitor + MGS fallback)."""
    n = V.shape[0]
    if k == 1:
        beta_prev, vprev = 0.0, np.zeros(n)           # decompositions.jl:64-74
    else:
        beta_prev, vprev = H[k - 2, k - 1], V[:, k - 2]   # decompositions.jl:76-83
    u = A @ V[:, k - 1]
    u = u - beta_prev * vprev
    H[k - 1, k - 1] = u @ V[:, k - 1]
    v = u - H[k - 1, k - 1] * V[:, k - 1]
    beta = float(np.linalg.norm(v))
    V[:, k] = 0.0 if beta == 0.0 else (1.0 / beta) * v
    if reorth:
        loss = orthogonality_loss(V, k + 1)
        if loss > SQRT_EPS:
            mgs_step(A, V, H, k)
            beta = H[k, k - 1]
            if k - 2 > 0:
                H[: k - 2, k - 1] = 0.0               # H[1:k-2, k] .= 0
            if stats is not None:
                stats["fallbacks"] = stats.get("fallbacks", 0) + 1
    H[k, k - 1] = beta                                # update_subdiagonals! decompositions.jl:180-186
    H[k - 1, k] = beta


# ----------------------------------------------------------------------------
# compressed solve (tensor_krylov_method.jl:10-34, utils.jl:501-546)
# ----------------------------------------------------------------------------
def _expm_times(Hk, gamma, sym_lower):
    """exp(gamma * first(H)):  Symmetric(:L) -> eigen path; Matrix -> Pade."""
    if sym_lower:
        S = np.tril(Hk) + np.tril(Hk, -1).T            # Symmetric(H, :L)  tensor_struct.jl:259
        w, Q = np.linalg.eigh(gamma * S)
        return (Q * np.exp(w)) @ Q.T
    M = gamma * Hk
    if np.array_equal(M, M.T):                         # Julia exp!(::Matrix) takes the eigen path when hermitian
        w, Q = np.linalg.eigh(M)
        return (Q * np.exp(w)) @ Q.T
    return sla.expm(M)


def solve_compressed_system(Hk_list, bt_list, alpha, omega, lambda_min, instance, cls, per_mode=False):
    """tensor_krylov_method.jl:10-34.  Returns (lambda, [Y_s k x t]).

    per_mode=False is the reference: exp(gamma * H_1) is applied to every mode
    (utils.jl:509-521).  EigValMat uses each mode's own H_s (utils.jl:525-546).
    per_mode=True is the mathematically intended variant (each mode its own H_s)."""
    d = len(Hk_list)
    k = Hk_list[0].shape[0]
    t = len(omega)
    lam_inv = 1.0 / lambda_min
    lam = lam_inv * np.asarray(omega)
    Y = [np.ones((k, t)) for _ in range(d)]
    own = per_mode or cls == EIGVALMAT
    for j in range(t):
        gamma = -alpha[j] * lam_inv
        if not own:
            E = _expm_times(Hk_list[0], gamma, instance == SYM)
            for s in range(d):
                Y[s][:, j] = E @ bt_list[s]
        else:
            for s in range(d):
                if cls == EIGVALMAT and not per_mode:
                    E = _expm_times(Hk_list[s], gamma, False)      # exp(gamma .* A[s]) on the raw view
                else:
                    E = _expm_times(Hk_list[s], gamma, instance == SYM)
                Y[s][:, j] = E @ bt_list[s]
    return lam, Y


def solve_compressed_fast(Hk_list, bt_list, alpha, omega, lambda_min, instance, per_mode=False):
    """Same result as solve_compressed_system for the symmetric eigen path, with ONE
    eigendecomposition per distinct matrix instead of t (flavour B of BASELINE.md)."""
    d = len(Hk_list)
    lam_inv = 1.0 / lambda_min
    lam = lam_inv * np.asarray(omega)
    gam = -np.asarray(alpha) * lam_inv
    Y = []
    Q = w = None
    for s in range(d):
        if per_mode or s == 0:
            Hk = Hk_list[s]
            if instance == SYM:
                S = np.tril(Hk) + np.tril(Hk, -1).T
                w, Q = np.linalg.eigh(S)
                Qi = Q.T
            else:
                w, Q = np.linalg.eig(Hk)
                Qi = np.linalg.inv(Q)
        c = Qi @ bt_list[s]
        Ys = Q @ (np.exp(np.outer(w, gam)) * c[:, None])
        Y.append(np.ascontiguousarray(Ys.real))
    return lam, Y


# ----------------------------------------------------------------------------
# residual estimate (utils.jl:132-443)
# ----------------------------------------------------------------------------
def _weights(t):
    return np.tril(2.0 * np.ones((t, t)), -1) + np.eye(t)


def gram_parts(Hk_list, Y, k):
    """Per-mode blocks used by the estimator: Ly, Z, X, Lz (utils.jl:186-204, 229-253, 285-288)."""
    Ly = [np.tril(Ys.T @ Ys) for Ys in Y]
    Z = [Hk_list[s] @ Y[s] for s in range(len(Y))]       # full H view, NOT the Symmetric wrapper (utils.jl:247)
    X = [Y[s].T @ Z[s] for s in range(len(Y))]
    Lz = [np.tril(Zs.T @ Zs) for Zs in Z]
    return Ly, Z, X, Lz


def residual_faithful(Hk_list, Y, lam, subdiag, bt_list, b_norm, k):
    """utils.jl:402-443 / 371-399 / 280-324 / 332-369 with the reference's loop
    structure (O(d^3 t^2)).  Returns dict(hy2, hyb, bb, boundary, r_comp, r_norm)."""
    d = len(Y)
    t = len(lam)
    Ly, Z, X, Lz = gram_parts(Hk_list, Y, k)
    Lam = np.tril(np.outer(lam, lam))                    # compute_lower_outer! :132-144
    W = _weights(t)
    boundary = 0.0
    for s in range(d):                                   # :428-437
        dl = Y[s][k - 1, :]
        Gam = np.tril(np.outer(dl, dl)) * Lam            # cp_tensor_coefficients :146-164
        P = np.ones((t, t))
        for r in range(d):
            if r != s:
                P = P * Ly[r]
        boundary += abs(subdiag[s]) ** 2 * float(np.sum(W * Gam * P))   # squared_tensor_entries :206-226
    hy2 = 0.0
    for s in range(d):                                   # MVnorm :296-318
        for r in range(d):
            P = np.ones((t, t))
            for q in range(d):
                if q != s and q != r:
                    P = P * Ly[q]
            term = P * Lz[s] if s == r else P * X[s] * X[r].T
            hy2 += float(np.sum(W * Lam * term))
    hyb = 0.0
    for s in range(d):                                   # tensorinnerprod :332-369 (first rows only)
        p = lam * Z[s][0, :]
        for q in range(d):
            if q != s:
                p = p * Y[q][0, :]
        hyb += float(p.sum())
    hyb *= b_norm
    bb = float(np.prod([bt @ bt for bt in bt_list]))     # kronproddot :392
    r_comp = hy2 - 2.0 * hyb + bb                        # :393
    out = dict(hy2=hy2, hyb=hyb, bb=bb, boundary=boundary, r_comp=r_comp)
    out["r_norm"] = math.sqrt(boundary + r_comp) if r_comp >= 0 and boundary + r_comp >= 0 else float("nan")
    return out


def residual_nilpotent(Hk_list, Y, lam, subdiag, bt_list, b_norm, k):
    """Same quantities in O(d t^2): product over modes in R[e,h]/(e^2,h^2)
    (SURVEY.md section 7).  Validated against residual_faithful in the tests."""
    d = len(Y)
    t = len(lam)
    Ly, Z, X, Lz = gram_parts(Hk_list, Y, k)
    Lam = np.tril(np.outer(lam, lam))
    W = _weights(t)
    P0 = np.ones((t, t)); Pe = np.zeros((t, t)); Ph = np.zeros((t, t)); Peh = np.zeros((t, t)); Pg = np.zeros((t, t))
    v0 = np.ones(t); v1 = np.zeros(t)
    for q in range(d):
        L = Ly[q]
        dl = Y[q][k - 1, :]
        g = abs(subdiag[q]) ** 2 * np.outer(dl, dl)
        Peh = Peh * L + Pe * X[q].T + Ph * X[q] + P0 * Lz[q]
        Pe = Pe * L + P0 * X[q]
        Ph = Ph * L + P0 * X[q].T
        Pg = Pg * L + P0 * g
        P0 = P0 * L
        v1 = v1 * Y[q][0, :] + v0 * Z[q][0, :]
        v0 = v0 * Y[q][0, :]
    hy2 = float(np.sum(W * Lam * np.tril(Peh)))
    boundary = float(np.sum(W * Lam * np.tril(Pg)))
    hyb = b_norm * float(np.sum(lam * v1))
    bb = float(np.prod([bt @ bt for bt in bt_list]))
    r_comp = hy2 - 2.0 * hyb + bb
    out = dict(hy2=hy2, hyb=hyb, bb=bb, boundary=boundary, r_comp=r_comp)
    out["r_norm"] = math.sqrt(boundary + r_comp) if r_comp >= 0 and boundary + r_comp >= 0 else float("nan")
    return out


# ----------------------------------------------------------------------------
# the driver (tensor_krylov_method.jl:36-125)
# ----------------------------------------------------------------------------
class OracleSolve:
    """Stateful restatement of tensorkrylov! so tests can inspect every iteration.

    A_list : d operators (scipy CSR or dense); b_list : d right-hand sides (already
    normalised if the caller went through TensorizedSystem, system.jl:33-37).
    """

    def __init__(self, A_list, b_list, tol, nmax, variant, instance, cls, tables=None,
                 per_mode=False, residual="nilpotent", fast_solve=False, schedule=None,
                 ignore_breakdown=False, mode_threads=1):
        self.A, self.b = list(A_list), [np.asarray(b, dtype=np.float64) for b in b_list]
        self.d = len(self.A)
        self.n = self.A[0].shape[0]
        self.tol, self.nmax, self.variant, self.instance, self.cls = tol, nmax, variant, instance, cls
        self.per_mode, self.fast_solve = per_mode, fast_solve
        self.ignore_breakdown = ignore_breakdown
        # the modes are independent inside a Krylov step: a thread pool over modes lets the CPU baseline use every
        # core (numpy/scipy release the GIL in the matvec, dot and Gram products); results do not depend on it
        self.pool = None
        if mode_threads > 1:
            from concurrent.futures import ThreadPoolExecutor
            self.pool = ThreadPoolExecutor(max_workers=mode_threads)
        self.residual_fn = residual_faithful if residual == "faithful" else residual_nilpotent
        self.schedule = schedule if schedule is not None else build_schedule(
            self.A[0], self.d, nmax, tol, instance, cls, tables)
        cols = nmax + 1
        self.V = [np.zeros((self.n, cols)) for _ in range(self.d)]          # decompositions.jl:130
        self.H = [np.zeros((cols, cols)) for _ in range(self.d)]            # decompositions.jl:131
        self.bt = [np.zeros(cols) for _ in range(self.d)]
        self.b_norm = math.sqrt(float(np.prod([b @ b for b in self.b])))    # kronprodnorm :48
        # ConvergenceData(nmax): everything = ones (convergence.jl:11-20)
        self.relres = np.ones(nmax); self.projres = np.ones(nmax); self.orth = np.ones(nmax)
        self.niterations = nmax
        self.status = None
        self.x = None
        self.detail = {}
        self.stats = {}
        self.k = 1
        def first(s):                                                       # :53, orthogonal_bases.jl:142-160
            self.V[s][:, 0] = (1.0 / np.linalg.norm(self.b[s])) * self.b[s] # initialize_decomp! decompositions.jl:112-118
            self._step(s, 1)
            self.bt[s][0] = self.V[s][:, 0] @ self.b[s]                     # initialize_compressed_rhs utils.jl:456-464
        self._over_modes(first)

    def _over_modes(self, fn):
        if self.pool is None:
            for s in range(self.d):
                fn(s)
        else:
            list(self.pool.map(fn, range(self.d)))

    def _step(self, s, k):
        if self.variant == ARNOLDI:
            mgs_step(self.A[s], self.V[s], self.H[s], k)
        else:
            lanczos_step(self.A[s], self.V[s], self.H[s], k, self.variant == LANCZOS_REORTH, self.stats)

    def iterate(self):
        """One pass of the loop body :63-120.  Returns the status if the solve ended, else None."""
        k = self.k + 1
        self.k = k
        d = self.d
        def advance(s):
            self._step(s, k)                                                # :66
            self.bt[s][k - 1] = self.V[s][:, k - 1] @ self.b[s]             # update_rhs! :71
        self._over_modes(advance)
        sc = self.schedule[k]
        Hk = [self.H[s][:k, :k] for s in range(d)]
        btk = [self.bt[s][:k] for s in range(d)]
        if self.fast_solve and self.cls != EIGVALMAT:
            lam, Y = solve_compressed_fast(Hk, btk, sc["alpha"], sc["omega"], sc["lambda_min"], self.instance,
                                           self.per_mode)
        else:
            lam, Y = solve_compressed_system(Hk, btk, sc["alpha"], sc["omega"], sc["lambda_min"], self.instance,
                                             self.cls, self.per_mode)         # :76
        sub = [self.H[s][k, k - 1] for s in range(d)]                       # :79
        r = self.residual_fn(Hk, Y, lam, sub, btk, self.b_norm, k)          # :83
        r["t"] = sc["t"]; r["lambda_min"] = sc["lambda_min"]
        self.detail[k] = r
        self.lastY, self.lastlam = Y, lam
        if r["r_comp"] < 0.0 and not self.ignore_breakdown:                 # utils.jl:395 -> :85-96
            self.niterations = k - 1
            self.relres = self.relres[: k - 1]; self.projres = self.projres[: k - 1]; self.orth = self.orth[: k - 1]
            self.status = ST_BREAKDOWN
            return self.status
        if r["r_comp"] < 0.0:
            r["r_norm"] = math.sqrt(max(r["boundary"] + r["r_comp"], 0.0))
        rel = r["r_norm"] / self.b_norm                                     # :99
        self.relres[k - 1] = rel
        self.projres[k - 1] = r["r_comp"]
        self.orth[k - 1] = orthogonality_loss(self.V[0], k)                 # :103
        if rel < self.tol and not self.ignore_breakdown:                    # :108-118
            self.x = (lam.copy(), [self.V[s][:, :k] @ Y[s] for s in range(d)])  # basis_tensor_mul! utils.jl:478-488
            self.status = ST_CONVERGED
            return self.status
        if k == self.nmax:
            self.status = ST_NMAX                                           # :122
            return self.status
        return None

    def run(self):
        while self.status is None:
            if self.nmax < 2:
                self.status = ST_NMAX
                break
            self.iterate()
        return self


def tensorkrylov(A_list, b_list, tol, nmax, variant, instance, cls, tables=None, **kw):
    return OracleSolve(A_list, b_list, tol, nmax, variant, instance, cls, tables, **kw).run()


# ----------------------------------------------------------------------------
# dense Kronecker oracle for tiny cases (the reference's test/utils.jl:88-185 idea)
# ----------------------------------------------------------------------------
def kron_sum_dense(mats):
    """A = sum_s I x ... x A_s x ... x I with mode 1 fastest (kroneckervectorize, tensor_struct.jl:361-384)."""
    sizes = [m.shape[0] for m in mats]
    N = int(np.prod(sizes))
    A = np.zeros((N, N))
    for s, M in enumerate(mats):
        M = M.toarray() if sp.issparse(M) else np.asarray(M)
        left = int(np.prod(sizes[s + 1:]))   # slower modes
        right = int(np.prod(sizes[:s]))      # faster modes
        A += np.kron(np.kron(np.eye(left), M), np.eye(right))
    return A


def kruskal_vectorize(lam, fmat):
    """tensor_struct.jl:361-384: sum_i lam_i * (x_d(:,i) kron ... kron x_1(:,i))."""
    N = int(np.prod([F.shape[0] for F in fmat]))
    out = np.zeros(N)
    for i in range(len(lam)):
        tmp = fmat[-1][:, i]
        for j in range(len(fmat) - 2, -1, -1):
            tmp = np.kron(tmp, fmat[j][:, i])
        out += lam[i] * tmp
    return out


def kron_vector(vs):
    tmp = vs[-1]
    for j in range(len(vs) - 2, -1, -1):
        tmp = np.kron(tmp, vs[j])
    return tmp
