"""Pins the CPU oracle against EVERY reproducible stored run of the reference:
27 of the 30 files under experiments/data -- Laplace / ConvDiff reproductions (two generations of the code),
smooth and non-smooth right-hand sides, the parameterised tridiagonal (alpha) and convection (beta) families, the
clustered-spectrum EigValMat runs with and without the dense epsilon-perturbation (a different dense operator in
every mode), and the uniform-spectrum run (a different diagonal operator in every mode).

The full sweep (all d, 40 iterations) is profiles/r01_corpus_oracle_report.txt; here d = 5 and 10 keep the CPU
suite short.  Tolerance model of SURVEY.md 8c: r_comp agrees to 1e-11 of the magnitude of the terms it cancels in
every run; the relative residual agrees to 1e-9 wherever it is not itself a cancellation result."""
import numpy as np
import pytest

import corpus as C


def _run(orc, tables, e, d, K):
    exp = dict(instance=e["instance"], orth=e["orth"])
    S = C.corpus_sweep.run_oracle(orc, tables, exp, dict(d=d, rhs=e["rhs"]), e["recipe"], K)
    k = np.arange(2, K + 1)
    mag = np.array([S.detail[kk]["hy2"] + 2 * abs(S.detail[kk]["hyb"]) + S.detail[kk]["bb"] for kk in k])
    return S, k, mag


@pytest.mark.parametrize("d", [5, 10])
@pytest.mark.parametrize("key", C.files())
def test_oracle_reproduces_stored_run(orc, tables, key, d):
    e = C.entry(key, d)
    K = min(24, e["length"])
    if K < 2:
        pytest.skip("the stored run ended before k = 2")
    S, k, mag = _run(orc, tables, e, d, K)
    rr, pr = e["relres"], e["projres"]
    e_comp = np.abs(S.projres[k - 1] - pr[k - 1]) / mag
    assert e_comp.max() < 1e-11, (key, d, e_comp.max())
    # relres^2 = (boundary + r_comp) / ||b||^2 with ||b|| = 1: same absolute bar on the squares
    assert np.max(np.abs(S.relres[k - 1] ** 2 - rr[k - 1] ** 2) / mag) < 1e-11
    well = rr[k - 1] ** 2 > 1e-3 * mag          # not dominated by the cancellation in r_comp
    if well.any():
        e_rel = np.abs(S.relres[k - 1] - rr[k - 1])[well] / rr[k - 1][well]
        assert e_rel.max() < 1e-9, (key, d, e_rel.max())


def test_corpus_covers_every_reproducible_file():
    keys = C.files()
    assert len(keys) == 27
    z = C.corpus()
    kinds = {str(z[f"{k}__meta"][3]) for k in keys}
    assert kinds == {"laplace", "convdiff", "sym_alpha", "nonsym_beta", "eig_zero", "eig_one", "eig_zero_eps",
                     "eig_one_eps", "eig_uniform"}
    for k in keys:
        assert list(z[f"{k}__dims"]) == [5, 10, 50, 100]


@pytest.mark.parametrize("kind,param", [("sym_alpha", (1.999756,)), ("nonsym_beta", (-3.0,)), ("nonsym_beta", (-5.07,)),
                                        ("eig_one_eps", (1e-2,)), ("eig_uniform", (1e-3, 1.0))])
def test_host_schedule_matches_oracle_on_corpus_operators(tk, orc, tables, kind, param):
    """The product's host side (api.extreme_eigvals + the library's table lookup / sinc coefficients) builds the
    same exp-sum schedule as the oracle for the operator families only the corpus reaches."""
    d, K = 10, 24
    A, ocls = C.corpus_sweep.operators(orc, (kind,) + param, d)
    nonsym = kind == "nonsym_beta"
    inst, oinst = (tk.NonSymInstance, orc.NONSYM) if nonsym else (tk.SymInstance, orc.SYM)
    cls = {orc.RANDSPD: tk.RandSPD, orc.CONVDIFF: tk.ConvDiff, orc.EIGVALMAT: tk.EigValMat}[ocls]
    ref = orc.build_schedule(A[0], d, K, 1e-9, oinst, ocls, None if nonsym else tables)
    for k in range(2, K + 1):
        lmin, lmax = tk.extreme_eigvals(A[0], d, k, inst, cls)
        assert lmin == ref[k]["lambda_min"]
        if nonsym:
            _, om, al = tk.nonsym_coefficients(lmin, 1e-9)
        else:
            t, om, al, _, _ = tk.sym_lookup(lmax * (1.0 / lmin), 1e-9)
            assert t == ref[k]["t"]
        np.testing.assert_array_equal(om, ref[k]["omega"])
        np.testing.assert_array_equal(al, ref[k]["alpha"])
