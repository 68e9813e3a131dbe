"""GPU parity at BASELINE.json's sizes: the CUDA path (through the C-ABI) against the CPU oracle.

The reference stores Julia results only for n = 200, so at n = 10^3..10^4 the pinned oracle is the available
reference.  The fixtures tests/golden/c{2,3,4,5}_oracle.npz are oracle runs on the bench's own inputs
(tools/make_parity_fixtures.py; tests/test_oracle_baseline_sizes.py re-derives them on the CPU):

  C2  d=50,   n=1000,  Laplace,  TensorLanczosReorth, nmax=256   whole solve, parity mode (exits, histories)
  C4  d=100,  n=2000,  ConvDiff, TensorArnoldi,       nmax=120   whole solve, parity mode
  C3  d=256,  n=10^4,  Laplace,  TensorLanczosReorth  first 16 iterations of the benchmarked solve
  C5  d=1024, n=10^4,  Laplace,  TensorLanczosReorth  first 16 iterations of the benchmarked solve

Tolerances: ||Hy||^2, <Hy,b>, ||b~||^2, Krylov coefficients and b~ to 1e-11 relative; r_comp to
1e-11 of the magnitude of the three terms it cancels; relres^2 to 4e-11 ||b||^2.  The boundary term is built from
the LAST rows of the Y_s, entries that decay to rounding level of the first rows within a few iterations (at C5,
k = 16 it is below 1e-20 while r_comp is 6e-13): it is compared to 1e-10 relative where it matters and otherwise on
the scale it enters the residual with, relres^2 = (boundary + r_comp)/||b||^2, i.e. the absolute bound of r_comp.
"""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, golden

pytestmark = pytest.mark.gpu

RTOL = 1e-11


def inputs(tk, name):
    cases = {
        "c2": dict(d=50, n=1000, cls=tk.Laplace, variant=tk.TensorLanczosReorth, instance=tk.SymInstance, nmax=256, fixed=False),
        "c4": dict(d=100, n=2000, cls=tk.ConvDiff, variant=tk.TensorArnoldi, instance=tk.NonSymInstance, nmax=120, fixed=False),
        "c3": dict(d=256, n=10000, cls=tk.Laplace, variant=tk.TensorLanczosReorth, instance=tk.SymInstance, nmax=16, fixed=True),
        "c5": dict(d=1024, n=10000, cls=tk.Laplace, variant=tk.TensorLanczosReorth, instance=tk.SymInstance, nmax=16, fixed=True),
    }
    c = cases[name]
    b = np.random.default_rng(12345).random(c["n"])
    b = b * (1.0 / np.linalg.norm(b))             # TensorizedSystem, system.jl:33-37
    A1 = tk.assemble_matrix(c["n"], c["cls"])
    return c, A1, b


def compare(slv, res, ref, name, nmax_run):
    """Every iteration the fixture holds against the device's record of the same solve; returns the worst errors."""
    ks = ref["k"]
    det = slv.detail(int(ks[0]), int(ks[-1]))
    worst = {}
    for key in ("hy2", "hyb", "bb"):
        err = np.abs(det[key] - ref[key]) / np.abs(ref[key])
        worst[key] = float(err.max())
    scale = np.abs(ref["hy2"]) + 2 * np.abs(ref["hyb"]) + np.abs(ref["bb"])
    db = np.abs(det["boundary"] - ref["boundary"])
    worst["boundary"] = float(np.minimum(db / np.abs(ref["boundary"]) / 10.0, db / scale).max())   # < 1e-11 passes
    big = ref["boundary"] > 1e-9 * scale              # iterations where the boundary term is far above rounding level
    worst["boundary_rel_first_iterations"] = float((db / np.abs(ref["boundary"]))[big].max()) if big.any() else 0.0
    worst["r_comp_over_terms"] = float((np.abs(det["r_comp"] - ref["r_comp"]) / scale).max())
    rr, rref = res["relres"][ks - 1], ref["relres"][ks - 1]
    worst["relres_sq_abs"] = float(np.abs(rr**2 - rref**2).max())
    well = rref > 1e-2                              # recorded, not asserted: implied by the bound on relres^2
    worst["relres_rel_where_well_conditioned"] = float((np.abs(rr - rref) / rref)[well].max()) if well.any() else 0.0
    kk = int(ks[-1])
    # columns 1..kk of H (column kk+1 belongs to a step a fixture cut at kk iterations never took)
    H = slv.get_H(0)[: kk + 1, :kk]
    worst["H1"] = float(np.abs(H - ref["H1"][:, :kk]).max() / np.abs(ref["H1"]).max())
    worst["bt1"] = float(np.abs(slv.get_bt(0)[: kk + 1] - ref["bt1"]).max() / np.abs(ref["bt1"]).max())
    worst["t_equal"] = bool(np.array_equal(det["t"].astype(int), ref["t"]))
    worst["lambda_min"] = float((np.abs(det["lambda_min"] - ref["lambda_min"]) / ref["lambda_min"]).max())
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"parity_{name}.json"), "w") as f:
        json.dump({"case": name, "iterations": [int(ks[0]), int(ks[-1])], "status": int(res["status"]), "worst": worst}, f)
    assert worst["t_equal"] and worst["lambda_min"] < 1e-14
    for key in ("hy2", "hyb", "bb", "H1", "bt1"):
        assert worst[key] < RTOL, (key, worst[key])
    assert worst["boundary"] < RTOL, worst
    assert worst["boundary_rel_first_iterations"] < 1e-10, worst
    assert worst["r_comp_over_terms"] < RTOL, worst["r_comp_over_terms"]
    assert worst["relres_sq_abs"] < 4 * RTOL, worst["relres_sq_abs"]
    return worst


@pytest.mark.parametrize("name", ["c2", "c4"])
def test_whole_solve_matches_oracle(tk, gpu, name):
    """C2 and C4 as BASELINE.json states them, parity mode: same exit, same iteration count, same histories."""
    c, A1, b = inputs(tk, name)
    ref = golden(name + "_oracle")
    slv = tk.Solver(c["d"], c["n"], c["nmax"], c["instance"], c["cls"], c["variant"], flags=tk.TK_FLAG_REFERENCE_H1)
    slv.set_operators([A1] * c["d"])
    slv.set_rhs([b] * c["d"])
    slv.set_schedule(A1, 1e-8)
    res = slv.solve(1e-8)
    assert res["status"] == int(ref["status"]) and res["niterations"] == int(ref["niterations"])
    assert res["term_k"] == int(ref["k"][-1])
    compare(slv, res, ref, name, c["nmax"])
    # the orthogonality loss is an accumulation of rounding errors: same size, not the same digits
    go, ro = res["orth"][1:], ref["orth"][1:]
    assert np.all(go <= 10 * ro + 1e-14) and np.all(ro <= 10 * go + 1e-14) and go.max() < 1.5e-8
    assert res["relres"][0] == 1.0 and res["projres"][0] == 1.0           # convergence.jl:11-20
    # a second solve on the handle replays the recorded CUDA graphs: bit-identical histories
    res2 = slv.solve(1e-8)
    assert slv.solve_info()["graphs_launched"] > 0
    assert np.array_equal(res2["relres"], res["relres"]) and np.array_equal(res2["projres"], res["projres"])
    slv.close()


@pytest.mark.parametrize("name", ["c3", "c5"])
def test_first_iterations_of_the_benchmarked_solve_match_oracle(tk, gpu, name):
    """C3 / C5 exactly as bench.py runs them (fixed-iteration mode, nmax = 64): iterations 2..16 against the oracle."""
    c, A1, b = inputs(tk, name)
    ref = golden(name + "_oracle")
    nmax = 64
    slv = tk.Solver(c["d"], c["n"], nmax, c["instance"], c["cls"], c["variant"],
                    flags=tk.TK_FLAG_REFERENCE_H1 | tk.TK_FLAG_FIXED_ITERATIONS)
    slv.set_operators([A1] * c["d"])
    slv.set_rhs([b] * c["d"])
    slv.set_schedule(A1, 1e-8)
    res = slv.solve(1e-8)
    assert res["status"] == tk.TK_NMAX and res["term_k"] == nmax
    compare(slv, res, ref, name, nmax)
    slv.close()


def test_converged_exit_returns_the_oracles_kruskal_tensor(tk, orc, tables, gpu):
    """d=256, n=10^4, tol 1e-5 converges at k=5 while later iterations are already enqueued,
    and t(k) changes right behind the exit (t(5)=5, t(7)=6): rank, lambda and factor matrices must be the ones of
    the iteration the loop left at.  Compared with the oracle's x (basis_tensor_mul!, utils.jl:478-488)."""
    d, n, nmax, tol = 256, 10000, 64, 1e-5
    b1 = np.random.default_rng(12345).random(n)
    A = tk.KroneckerMatrix.gallery(tk.SymInstance, d, n, tk.Laplace)
    system = tk.TensorizedSystem(tk.SymInstance, A, [b1] * d)
    cd = tk.ConvergenceData(nmax)
    keep = []
    x = tk.tensorkrylov(cd, system.A, system.b, tol, nmax, tk.TensorLanczosReorth, verbose=False, solver_out=keep)
    Ao = orc.assemble_matrix(n, orc.LAPLACE)
    S = orc.tensorkrylov([Ao] * d, orc.normalize_rhs([b1] * d), tol, nmax, orc.LANCZOS_REORTH, orc.SYM, orc.LAPLACE,
                         tables, mode_threads=os.cpu_count() or 1)
    assert S.status == orc.ST_CONVERGED
    assert cd.status == tk.TK_CONVERGED and x is not None
    assert cd.term_k == S.k and cd.niterations == nmax
    lam, fm = S.x
    assert x.ncomponents() == len(lam) == S.schedule[S.k]["t"]
    assert np.abs(x.lambda_ - lam).max() <= 1e-13 * np.abs(lam).max()
    for s in (0, 1, d // 2, d - 1):                          # includes modes s > 0: a wrong rank corrupts exactly those
        assert x.fmat[s].shape == fm[s].shape
        assert np.abs(x.fmat[s] - fm[s]).max() <= 1e-9 * np.abs(fm[s]).max()
    kk = np.arange(2, S.k + 1)
    assert np.abs(cd.relative_residual_norm[kk - 1] ** 2 - S.relres[kk - 1] ** 2).max() <= 4 * RTOL
    slv = keep[0]
    # the single-mode getter, the pinned destination and a too-small buffer
    lam1, F1 = slv.solution_mode(d - 1)
    assert np.array_equal(F1, x.fmat[d - 1]) and np.array_equal(lam1, x.lambda_)
    lam2, fp = slv.solution(pinned=True)
    assert all(np.array_equal(fp[s], x.fmat[s]) for s in (0, 7, d - 1))
    import ctypes as C
    small = np.zeros(4)
    rc = tk._capi.lib.tk_get_solution_all(slv.h, tk._capi.dptr(small), 1, tk._capi.dptr(small), 4, 0)
    assert rc == -1 and b"rank" in tk._capi.lib.tk_last_error()
    slv.close()


def test_true_residual_after_an_early_converged_exit(tk, orc, tables, gpu):
    """Small dense Kronecker check of the same path: a solve that converges early returns an x whose TRUE residual
    ||Ax - b|| / ||b|| equals the estimator's value at the exit iteration."""
    d, n, nmax = 3, 14, 13
    rng = np.random.default_rng(5)
    A1 = tk.assemble_matrix(n, tk.Laplace)
    b = orc.normalize_rhs([rng.random(n) for _ in range(d)])
    ref = orc.tensorkrylov([orc.assemble_matrix(n, orc.LAPLACE)] * d, b, 1e-8, nmax, orc.LANCZOS_REORTH, orc.SYM,
                           orc.LAPLACE, tables, per_mode=True, ignore_breakdown=True)
    # a tolerance the run crosses three iterations before nmax (the first crossing may come earlier)
    k_exit = nmax - 3
    tol = float(ref.relres[k_exit - 1]) * 1.0000001
    slv = tk.Solver(d, n, nmax, tk.SymInstance, tk.Laplace, tk.TensorLanczosReorth, flags=0)
    slv.set_operators([A1] * d)
    slv.set_rhs(b)
    slv.set_schedule(A1, 1e-8)
    res = slv.solve(tol)
    assert res["status"] == tk.TK_CONVERGED and 2 <= res["term_k"] <= k_exit < nmax
    lam, fmat = slv.solution()
    x = tk.kroneckervectorize(tk.KruskalTensor(lam, [fmat[s] for s in range(d)]))
    Ad = orc.kron_sum_dense([A1] * d)
    bd = orc.kron_vector(b)
    true = np.linalg.norm(Ad @ x - bd) / np.linalg.norm(bd)
    assert res["relres"][res["term_k"] - 1] == pytest.approx(true, rel=1e-6)
    slv.close()


def test_library_schedule_matches_lapack_schedule(tk, gpu):
    """f1: tk_schedule (eigen-extremes of the minors of A_1 computed inside the library) against the schedule fed from
    LAPACK through tk_set_schedule, on the C4 operator: same ranks, lambda_min to eps * cond(minor), same histories
    to 1e-9."""
    d, n, nmax = 8, 2000, 60
    b = np.random.default_rng(12345).random(n)
    b = b * (1.0 / np.linalg.norm(b))
    A1 = tk.assemble_matrix(n, tk.ConvDiff)
    out = {}
    for spectral in ("library", "lapack"):
        slv = tk.Solver(d, n, nmax, tk.NonSymInstance, tk.ConvDiff, tk.TensorArnoldi,
                        flags=tk.TK_FLAG_REFERENCE_H1 | tk.TK_FLAG_FIXED_ITERATIONS)
        slv.set_operators([A1] * d)
        slv.set_rhs([b] * d)
        slv.set_schedule(A1, 1e-8, spectral=spectral)
        res = slv.solve(1e-8)
        out[spectral] = (res, slv.detail())
        slv.close()
    (ra, da), (rb, db) = out["library"], out["lapack"]
    assert np.array_equal(da["t"], db["t"])
    assert np.max(np.abs(da["lambda_min"] - db["lambda_min"]) / db["lambda_min"]) < 1e-11
    assert np.max(np.abs(ra["relres"] - rb["relres"]) / rb["relres"]) < 1e-9
    # SymInstance classes: EigValMat (diagonal extremes) and a tridiagonal operator under the RandSPD rule
    ev = (np.arange(1, 201) / 200.0) ** 2
    for cls, A in ((tk.EigValMat, tk.assemble_matrix(ev, tk.EigValMat)),
                   (tk.RandSPD, np.diag(2.1 * np.ones(200)) - np.diag(np.ones(199), 1) - np.diag(np.ones(199), -1))):
        hist = {}
        for spectral in ("library", "lapack"):
            slv = tk.Solver(3, 200, 40, tk.SymInstance, cls, tk.TensorLanczosReorth,
                            flags=tk.TK_FLAG_REFERENCE_H1 | tk.TK_FLAG_FIXED_ITERATIONS)
            bb = np.random.default_rng(7).random(200)
            bb /= np.linalg.norm(bb)
            slv.set_operators([A] * 3)
            slv.set_rhs([bb] * 3)
            slv.set_schedule(A, 1e-8, spectral=spectral)
            hist[spectral] = (slv.solve(1e-8), slv.detail())
            slv.close()
        assert np.array_equal(hist["library"][1]["t"], hist["lapack"][1]["t"])
        assert np.max(np.abs(hist["library"][1]["lambda_min"] - hist["lapack"][1]["lambda_min"])
                      / hist["lapack"][1]["lambda_min"]) < 1e-10
