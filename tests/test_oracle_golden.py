"""Pins the CPU oracle against the reference's own stored results and known-answer tests.

Golden data: tests/golden/*.npz, decoded from /root/reference/experiments/data/** by
tools/make_golden.py (Julia-serialized `Experiment` objects: the right-hand sides the authors
used and the histories their Float64 Julia run produced).
"""
import numpy as np
import pytest

from conftest import golden


def test_laplace_history_d5(orc, tables):
    """reproduction_data/laplace_new, d=5: Laplace n=200, TensorLanczosReorth, tol 1e-9, nmax 199.
    The stored relative residual is reproduced to 1e-11 while r_comp is well above its
    cancellation noise (k <= 60), as established in SURVEY.md section 4."""
    g = golden("laplace_new")
    d, n, kmax = 5, 200, 60
    A = orc.assemble_matrix(n, orc.LAPLACE)
    S = orc.OracleSolve([A] * d, [g["rhs_d5"]] * d, 1e-9, 199, orc.LANCZOS_REORTH, orc.SYM, orc.LAPLACE, tables,
                        residual="faithful")
    while S.k < kmax:
        assert S.iterate() is None
    rr, pr = g["relres_d5"], g["projres_d5"]
    k = np.arange(2, kmax + 1)
    assert np.max(np.abs(S.relres[k - 1] - rr[k - 1]) / rr[k - 1]) < 1e-11
    assert np.max(np.abs(S.projres[k - 1] - pr[k - 1]) / pr[k - 1]) < 1e-10
    # exp-sum ranks follow the CSV as shipped (t = 6 at k=2, 7 at k=3, 15 at k=30)
    assert [S.detail[kk]["t"] for kk in (2, 3, 30)] == [6, 7, 15]
    assert rr[1] == pytest.approx(7.2179574187e-01, rel=1e-10)


def test_laplace_history_d10_nilpotent(orc, tables):
    """Same experiment, d=10, through the O(d t^2) combine -- validates the nilpotent-algebra restatement
    against the reference's stored numbers, not only against the brute-force loops."""
    g = golden("laplace_new")
    d, n, kmax = 10, 200, 40
    A = orc.assemble_matrix(n, orc.LAPLACE)
    S = orc.OracleSolve([A] * d, [g["rhs_d10"]] * d, 1e-9, 199, orc.LANCZOS_REORTH, orc.SYM, orc.LAPLACE, tables,
                        residual="nilpotent", fast_solve=True)
    while S.k < kmax:
        assert S.iterate() is None
    rr = g["relres_d10"]
    k = np.arange(2, kmax + 1)
    assert np.max(np.abs(S.relres[k - 1] - rr[k - 1]) / rr[k - 1]) < 1e-10


def test_convdiff_history_d5(orc):
    """reproduction_data/nonsym_new, d=5: ConvDiff n=200, TensorArnoldi (Pade exponential of H_1)."""
    g = golden("nonsym_new")
    d, n, kmax = 5, 200, 30
    A = orc.assemble_matrix(n, orc.CONVDIFF)
    S = orc.OracleSolve([A] * d, [g["rhs_d5"]] * d, 1e-9, 200, orc.ARNOLDI, orc.NONSYM, orc.CONVDIFF, None)
    while S.k < kmax:
        assert S.iterate() is None
    rr = g["relres_d5"]
    k = np.arange(2, kmax + 1)
    assert np.max(np.abs(S.relres[k - 1] - rr[k - 1]) / rr[k - 1]) < 1e-11
    assert S.detail[2]["t"] == 39 and abs(S.detail[2]["lambda_min"] - 2.046548e5) < 1.0


def test_eigvalmat_history_d5(orc, tables):
    """eigenvalues_data/dzero, d=5: EigValMat with eigenvalues j^2/n^2 -- exercises the EigValMat spectral rule
    (eigenvalues.jl:339) and the per-mode exp(gamma .* H_s) variant (utils.jl:525-546)."""
    g = golden("eigval_dzero")
    d, n, kmax = 5, 200, 40
    # clusterzero (experiments/eigenvalue_distribution.jl:110-116): j^2 * inv(n^2) -- the multiplication by the
    # reciprocal matters: kappa = k^2 lands on either side of a table row boundary depending on the last bit
    ev = np.array([(j * j) * (1.0 / (n * n)) for j in range(1, n + 1)], dtype=np.float64)
    A = orc.assemble_matrix(n, orc.EIGVALMAT, eigenvalues=ev)
    S = orc.OracleSolve([A] * d, [g["rhs_d5"]] * d, 1e-9, 200, orc.LANCZOS_REORTH, orc.SYM, orc.EIGVALMAT, tables)
    while S.k < kmax:
        assert S.iterate() is None
    rr = g["relres_d5"]
    k = np.arange(2, kmax + 1)
    assert np.max(np.abs(S.relres[k - 1] - rr[k - 1]) / rr[k - 1]) < 1e-9


def test_known_answer_squared_tensor_entries(orc):
    """test/utils.jl:188-227: Y1=[2 1;1 2], Y2=[3 4;3 4], Y3=[2 2;2 2], lambda=1 -> [1936, 1768, 1768]."""
    Y = [np.array([[2.0, 1.0], [1.0, 2.0]]), np.array([[3.0, 4.0], [3.0, 4.0]]), np.array([[2.0, 2.0], [2.0, 2.0]])]
    lam = np.ones(2)
    Ly = [np.tril(y.T @ y) for y in Y]
    Lam = np.tril(np.outer(lam, lam))
    W = np.tril(2 * np.ones((2, 2)), -1) + np.eye(2)
    got = []
    for s in range(3):
        dl = Y[s][1, :]
        Gam = np.tril(np.outer(dl, dl)) * Lam
        P = np.ones((2, 2))
        for r in range(3):
            if r != s:
                P = P * Ly[r]
        got.append(float(np.sum(W * Gam * P)))
    assert got == [1936.0, 1768.0, 1768.0]
    # and through the oracle's boundary term with unit sub-diagonals
    H = [np.zeros((2, 2))] * 3
    r = orc.residual_faithful(H, Y, lam, [1.0, 1.0, 1.0], [np.zeros(2)] * 3, 1.0, 2)
    assert r["boundary"] == 1936.0 + 1768.0 + 1768.0
    r2 = orc.residual_nilpotent(H, Y, lam, [1.0, 1.0, 1.0], [np.zeros(2)] * 3, 1.0, 2)
    assert r2["boundary"] == pytest.approx(r["boundary"], rel=1e-15)


def test_known_answer_lanczos_sturm_values(orc):
    """test/eigenvalues.jl:5-41: 50x50 tridiag(-1,2,-1), v = ones/sqrt(n), 5 Lanczos steps; the characteristic
    polynomials of the leading minors of H evaluated at mu = 2."""
    n, k = 50, 5
    import scipy.sparse as sp
    A = sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1], format="csr")
    V = np.zeros((n, k + 1)); H = np.zeros((k + 1, k + 1))
    V[:, 0] = np.ones(n) / np.sqrt(n)
    for j in range(1, k + 1):
        orc.lanczos_step(A, V, H, j, reorth=False)
    mu = 2.0
    p = [1.0, H[0, 0] - mu]
    for j in range(2, k + 1):
        p.append((H[j - 1, j - 1] - mu) * p[-1] - H[j - 1, j - 2] ** 2 * p[-2])
    want = [1.0, -1.96, -0.04166666666666724, 1.9565217391304344, 0.04545454545454549, -1.9523809523809523]
    assert np.allclose(p, want, rtol=1e-10, atol=0)
    exact = np.linalg.eigvalsh(H[:k, :k])
    pk = np.poly(H[:k, :k])
    assert np.all(np.abs(np.polyval(pk, exact)) < 1e-13)


def test_analytic_eigenvalues_match_minors(orc):
    """test/eigenvalues.jl:75-100: analytic_eigenvalues(d,n,i) == d * extremes(eigvals(A[1:i,1:i]))."""
    d, n = 5, 200
    A = orc.assemble_matrix(n, orc.LAPLACE_DENSE)
    for i in [1, 2, 3, 10, 57, 199]:
        lmin, lmax = orc.laplace_extremes(d, n, i)
        ev = np.linalg.eigvalsh(A[:i, :i])
        assert lmin == pytest.approx(d * ev.min(), rel=1e-9)
        assert lmax == pytest.approx(d * ev.max(), rel=1e-12)


def test_residual_nilpotent_equals_faithful(orc, tables):
    """The O(d t^2) combine equals the reference's O(d^3 t^2) loops (distinct modes, d = 4)."""
    rng = np.random.default_rng(7)
    d, n, nmax = 4, 60, 12
    A = [orc.assemble_matrix(n, orc.LAPLACE)] * d
    b = orc.normalize_rhs([rng.random(n) for _ in range(d)])
    Sf = orc.tensorkrylov(A, b, 1e-8, nmax, orc.LANCZOS_REORTH, orc.SYM, orc.LAPLACE, tables, residual="faithful",
                          ignore_breakdown=True)
    Sn = orc.tensorkrylov(A, b, 1e-8, nmax, orc.LANCZOS_REORTH, orc.SYM, orc.LAPLACE, tables, residual="nilpotent",
                          ignore_breakdown=True)
    for k in range(2, nmax + 1):
        for key in ("hy2", "hyb", "bb", "boundary"):
            assert Sn.detail[k][key] == pytest.approx(Sf.detail[k][key], rel=1e-13)


def test_dense_kronecker_residual_identity(orc, tables):
    """The idea of test/utils.jl:88-185: for a tiny system the estimator's residual equals the true residual
    ||A x - b|| of the Kruskal iterate, computed with explicit Kronecker sums."""
    rng = np.random.default_rng(3)
    d, n, nmax = 3, 9, 6
    A1 = orc.assemble_matrix(n, orc.LAPLACE)
    b = orc.normalize_rhs([rng.random(n) for _ in range(d)])
    S = orc.OracleSolve([A1] * d, b, 1e-8, nmax, orc.LANCZOS_REORTH, orc.SYM, orc.LAPLACE, tables, per_mode=True,
                        residual="faithful", ignore_breakdown=True)
    Ad = orc.kron_sum_dense([A1] * d)
    bd = orc.kron_vector(b)
    for _ in range(2, nmax + 1):
        S.iterate()
        k = S.k
        x = orc.kruskal_vectorize(S.lastlam, [S.V[s][:, :k] @ S.lastY[s] for s in range(d)])
        true = np.linalg.norm(Ad @ x - bd) / np.linalg.norm(bd)
        assert S.relres[k - 1] == pytest.approx(true, rel=1e-6)


def test_oracle_mode_threads_do_not_change_results(orc, tables):
    """The CPU baseline runs the (independent) modes on a thread pool; the histories are bit-identical."""
    A = orc.assemble_matrix(300, orc.LAPLACE)
    b = orc.normalize_rhs(orc.random_rhs(12, 300))
    r1 = orc.tensorkrylov([A] * 12, b, 1e-8, 12, orc.LANCZOS_REORTH, orc.SYM, orc.LAPLACE, tables, ignore_breakdown=True)
    r4 = orc.tensorkrylov([A] * 12, b, 1e-8, 12, orc.LANCZOS_REORTH, orc.SYM, orc.LAPLACE, tables, ignore_breakdown=True,
                          mode_threads=4)
    assert np.array_equal(r1.relres, r4.relres) and np.array_equal(r1.projres, r4.projres)
