#!/usr/bin/env python3
"""Regression corpus over ALL stored runs of the reference.

`experiments/data/**` holds 30 Julia-serialized `Experiment` objects (experiments/experiment_common.jl:14-42):
the right-hand sides the authors drew and the `ConvergenceData` histories their Float64 Julia run produced for
d = 5, 10, 50, 100 at n = 200.  This script

  1. decodes every file (dims, class names, one rhs per d, iterations / relative residual / projected residual /
     orthogonality history per d);
  2. rebuilds the operators from the experiment drivers (reproduction.jl:10-21, eigenvalue_distribution.jl:110-133
     and 157-215, parameterized_systems.jl:3-23).  The drivers' free parameters (alpha, beta, epsilon, the
     eigenvalue interval) are not stored.  They were recovered once by matching the first iterations of the d=5
     history over successively finer grids (`--refit` repeats the last, narrow stage); every recovered value is a
     short decimal that reproduces the d=5 history to ~1e-14 while its grid neighbours are off by >= 1e-4, and it
     is then VERIFIED on d = 10, 50, 100, which played no part in the fit.  The `d2*/d5*/d14*` files are runs with
     epsilon = 1e-2 / 1e-5 / 1e-14 of an EARLIER version of `perturb_matrix!` (eigenvalue_distribution.jl:68-78:
     A[s] = (s*eps) .+ A[s], a dense rank-one shift of every entry, distinct per mode; the shipped driver only
     calls it when eps == 0, :92-96, which is a no-op);
  3. runs the CPU oracle on the same inputs and measures how far it is from the stored Julia numbers;
  4. writes tests/golden/corpus.npz (inputs + stored histories + recovered parameters; small) and a text report.

Runs in the build container only (needs /root/reference).  The tests read corpus.npz, never the reference.

Not reproducible and therefore listed but not compared: the two RandSPD runs (the random matrix is not stored)
and parametrized_data/uniform (no interval / construction in the shipped drivers reproduces it).
"""
import argparse
import os
import re
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import __graft_entry__ as entry               # noqa: E402

REF = os.environ.get("TK_REFERENCE", "/root/reference")
DATA = os.path.join(REF, "experiments", "data")
N = 200
TOL = 1e-9                                    # every driver's default (reproduction.jl:10, parameterized_systems.jl:56)

def decode(path):
    """One serialized Experiment through the product's own decoder (tensorkrylov.jl_b200/experiments.py)."""
    ex = entry.load_package().experiments.deserialize_from_file(path, N)
    runs = []
    for d, rhs, c in zip(ex.dims, ex.rhs_vec, ex.conv_vector):
        assert all(np.array_equal(rhs[0], r) for r in rhs)      # system.jl:5-11: one vector, d times
        runs.append(dict(d=d, rhs=rhs[0], iterations=c.iterations, relres=c.relative_residual_norm,
                         projres=c.projected_residual_norm, orth=c.orthogonality_data))
    return dict(instance=ex.instance.__name__, cls=ex.matrixclass.__name__, orth=ex.orth_method.__name__, runs=runs)


# ---- operators of the experiment drivers ----------------------------------------------------------------------
def sym_parameterized(alpha):
    """parameterized_systems.jl:3-10: inv(h^2) * SymTridiagonal(alpha ones(n), -ones(n-1))."""
    h = 1.0 / (N + 1)
    return (sp.diags([-np.ones(N - 1), alpha * np.ones(N), -np.ones(N - 1)], [-1, 0, 1], format="csr")
            * (1.0 / (h * h))).tocsr()


def nonsym_parameterized(beta):
    """parameterized_systems.jl:12-20: L + (10/(4h)) * diagm(-1=>1, 0=>3, 1=>beta, 2=>1)."""
    h = 1.0 / (N + 1)
    L = sp.diags([-np.ones(N - 1), 2.0 * np.ones(N), -np.ones(N - 1)], [-1, 0, 1], format="csr") * (1.0 / (h * h))
    C = sp.diags([np.ones(N - 1), 3.0 * np.ones(N), beta * np.ones(N - 1), np.ones(N - 2)], [-1, 0, 1, 2],
                 format="csr") * (10.0 * (1.0 / (4.0 * h)))
    return (L + C).tocsr()


def clusterzero():
    """eigenvalue_distribution.jl:110-116."""
    return np.array([(j * j) * (1.0 / (N * N)) for j in range(1, N + 1)])


def clusterone():
    """eigenvalue_distribution.jl:118-133."""
    v = np.zeros(N)
    v[0] = 1.0 / (N * N)
    tmp = np.log(float(N))
    for j in range(2, N + 1):
        v[j - 1] = np.log(float(j)) * (1.0 / tmp)
    return v


def uniform_modes(d, lo, hi):
    """eigenvalue_distribution.jl:157-173: mode s has the diagonal ((s-1) * step / d) .+ LinRange(lo, hi, n)."""
    ev = np.linspace(lo, hi, N)
    step = ev[1] - ev[0]
    return [((s * step) * (1.0 / d)) + ev for s in range(d)]


def operators(orc, recipe, d):
    """recipe = (kind, parameter...) -> (A_list, oracle class)."""
    kind = recipe[0]
    if kind == "laplace":
        return [orc.assemble_matrix(N, orc.LAPLACE)] * d, orc.LAPLACE
    if kind == "convdiff":
        return [orc.assemble_matrix(N, orc.CONVDIFF)] * d, orc.CONVDIFF
    if kind == "sym_alpha":                      # class tag RandSPD: extremes from eigvals of the minors
        return [sym_parameterized(recipe[1])] * d, orc.RANDSPD
    if kind == "nonsym_beta":
        return [nonsym_parameterized(recipe[1])] * d, orc.CONVDIFF
    if kind == "eig_zero":
        return [orc.assemble_matrix(N, orc.EIGVALMAT, eigenvalues=clusterzero())] * d, orc.EIGVALMAT
    if kind == "eig_one":
        return [orc.assemble_matrix(N, orc.EIGVALMAT, eigenvalues=clusterone())] * d, orc.EIGVALMAT
    if kind in ("eig_zero_eps", "eig_one_eps"):  # earlier perturb_matrix!: every entry of mode s shifted by s*eps
        ev = clusterzero() if kind == "eig_zero_eps" else clusterone()
        return [np.diag(ev) + (s * recipe[1]) for s in range(1, d + 1)], orc.EIGVALMAT
    if kind == "eig_uniform":
        return [np.diag(e) for e in uniform_modes(d, recipe[1], recipe[2])], orc.EIGVALMAT
    raise ValueError(kind)


VARIANT = {"TensorLanczos": 0, "TensorLanczosReorth": 1, "TensorArnoldi": 2}


def run_oracle(orc, tables, exp, run, recipe, kmax):
    d = run["d"]
    A, cls = operators(orc, recipe, d)
    inst = orc.NONSYM if exp["instance"] == "NonSymInstance" else orc.SYM
    b = orc.normalize_rhs([run["rhs"]] * d)
    S = orc.OracleSolve(A, b, TOL, kmax, VARIANT[exp["orth"]], inst, cls,
                        tables if inst == orc.SYM else None, ignore_breakdown=True,
                        fast_solve=(cls != orc.EIGVALMAT and inst == orc.SYM))
    while S.k < kmax:
        S.iterate()
    return S


def deviation(S, run, kmax):
    """max over k=2..kmax of: relative deviation of the relative residual; deviation of r_comp measured against
    the magnitude of the terms it cancels."""
    k = np.arange(2, kmax + 1)
    rr, pr = run["relres"], run["projres"]
    e_rel = np.abs(S.relres[k - 1] - rr[k - 1]) / np.abs(rr[k - 1])
    mag = np.array([S.detail[kk]["hy2"] + 2 * abs(S.detail[kk]["hyb"]) + S.detail[kk]["bb"] for kk in k])
    e_abs = np.abs(S.projres[k - 1] - pr[k - 1]) / mag
    return e_rel, e_abs


def recover(orc, tables, exp, kind, grid, kfit=6):
    """Pick the candidate whose d=5 history matches the stored one best over k=2..kfit.
    Returns (deviation of the best, its recipe, deviation of the runner-up)."""
    run = exp["runs"][0]
    kfit = min(kfit, len(run["iterations"]))
    best, scores = None, []
    for cand in grid:
        recipe = (kind,) + (cand if isinstance(cand, tuple) else (cand,))
        try:
            S = run_oracle(orc, tables, exp, run, recipe, kfit)
        except Exception:                         # indefinite minors, complex eigenvalues, kappa outside the table
            continue
        e_rel, _ = deviation(S, run, kfit)
        if not np.all(np.isfinite(e_rel)):
            continue
        score = float(e_rel.max())
        scores.append(score)
        if best is None or score < best[0]:
            best = (score, recipe)
    scores.sort()
    return best[0], best[1], (scores[1] if len(scores) > 1 else float("nan"))


SYM_ALPHA = {1: 1.9999, 2: 2.005, 3: 1.9998, 4: 1.99976, 5: 1.999756}
NONSYM_BETA = {1: -3.0, 2: -5.005, 3: -5.025, 4: -5.05, 5: -5.07}
EPS = {"d2": 1e-2, "d5": 1e-5, "d14": 1e-14}

FILES = [
    ("reproduction_data/laplace", ("laplace",)),
    ("reproduction_data/laplace_new", ("laplace",)),
    ("reproduction_data/nonsym", ("convdiff",)),
    ("reproduction_data/nonsym_new", ("convdiff",)),
    ("reproduction_data/rand_spd", None),
    ("reproduction_data/rand_spd_reorth", None),
    ("rhs_data/smooth", ("laplace",)),
    ("rhs_data/nonsmooth", ("laplace",)),
    ("parametrized_data/sym", ("laplace",)),
    ("parametrized_data/nonsym", ("convdiff",)),
] + [(f"parametrized_data/sym{i}", ("sym_alpha", a)) for i, a in SYM_ALPHA.items()] \
  + [(f"parametrized_data/nonsym{i}", ("nonsym_beta", b)) for i, b in NONSYM_BETA.items()] \
  + [("eigenvalues_data/dzero", ("eig_zero",)), ("eigenvalues_data/done", ("eig_one",))] \
  + [(f"eigenvalues_data/{p}zero", ("eig_zero_eps", e)) for p, e in EPS.items()] \
  + [(f"eigenvalues_data/{p}one", ("eig_one_eps", e)) for p, e in EPS.items()] \
  + [("eigenvalues_data/uniform", ("eig_uniform", 1e-3, 1.0)), ("parametrized_data/uniform", None)]


def refit_grid(recipe):
    """The last stage of the recovery: the recovered value and its neighbours one grid step away."""
    kind = recipe[0]
    if kind == "sym_alpha":
        return [round(recipe[1] + i * 1e-6, 6) for i in range(-3, 4)]
    if kind == "nonsym_beta":
        return [round(recipe[1] + i * 1e-3, 3) for i in range(-3, 4)]
    if kind in ("eig_zero_eps", "eig_one_eps"):
        return [recipe[1] * f for f in (0.1, 0.5, 1.0, 2.0, 10.0)]
    if kind == "eig_uniform":
        return [(recipe[1] * f, recipe[2]) for f in (0.5, 0.9, 1.0, 1.1, 2.0)]
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kmax", type=int, default=40, help="iterations compared per run")
    ap.add_argument("--refit", action="store_true", help="re-run the last stage of the parameter recovery")
    ap.add_argument("--report", default=os.path.join(ROOT, "profiles", "r01_corpus_oracle_report.txt"))
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "corpus.npz"))
    args = ap.parse_args()
    orc = entry.load_oracle()
    tables = orc.ExpSumTables.from_reference_dir(os.path.join(REF, "coefficients_data"))
    store, lines = {}, []
    lines.append("# oracle vs the reference's stored Julia histories, all files under experiments/data (n = 200, tol 1e-9)")
    lines.append("# e_rel: max_k |relres - stored| / stored;  e_comp: max_k |r_comp - stored| / (|Hy|^2 + 2|<Hy,b>| + |b|^2)")
    lines.append("# k<=K: iterations compared (stored history length or --kmax, whichever is shorter)")
    lines.append(f"{'file':34s} {'class':10s} {'variant':20s} {'recipe':28s} {'d':>4s} {'len':>4s} {'K':>3s} "
                 f"{'e_rel(k<=10)':>12s} {'e_rel(k<=K)':>12s} {'e_comp(k<=K)':>12s}")
    names = []
    for rel, recipe in FILES:
        exp = decode(os.path.join(DATA, rel))
        key = rel.replace("/", "__")
        if recipe is None:
            lines.append(f"{rel:34s} {exp['cls']:10s} {exp['orth']:20s} {'(inputs not stored)':28s}   -- not reproducible, "
                         f"history lengths {[len(r['iterations']) for r in exp['runs']]}")
            continue
        fit = ""
        grid = refit_grid(recipe) if args.refit else None
        if grid:
            score, refit, runner_up = recover(orc, tables, exp, recipe[0], grid)
            assert refit == recipe, (rel, refit, recipe)
            fit = f" (refit: d=5 k<=6 deviation {score:.1e}, best neighbour {runner_up:.1e})"
        names.append(key)
        store[f"{key}__meta"] = np.array([exp["instance"], exp["cls"], exp["orth"], recipe[0]])
        store[f"{key}__param"] = np.array(recipe[1:], dtype=np.float64)
        store[f"{key}__dims"] = np.array([r["d"] for r in exp["runs"]])
        for run in exp["runs"]:
            d = run["d"]
            for f in ("rhs", "iterations", "relres", "projres", "orth"):
                store[f"{key}__{f}_d{d}"] = run[f]
            K = min(args.kmax, len(run["iterations"]))
            if K < 2:
                lines.append(f"{rel:34s} {exp['cls']:10s} {exp['orth']:20s} {str(recipe):28s} {d:4d} {len(run['iterations']):4d}   -- "
                             f"stored run stopped before k=2")
                continue
            try:
                S = run_oracle(orc, tables, exp, run, recipe, K)
                e_rel, e_abs = deviation(S, run, K)
                k10 = min(K, 10) - 1
                lines.append(f"{rel:34s} {exp['cls']:10s} {exp['orth']:20s} {str(recipe):28s} {d:4d} {len(run['iterations']):4d} {K:3d} "
                             f"{e_rel[:k10].max():12.2e} {e_rel.max():12.2e} {e_abs.max():12.2e}{fit}")
            except Exception as e:                # noqa: BLE001 -- report, do not hide
                lines.append(f"{rel:34s} {exp['cls']:10s} {exp['orth']:20s} {str(recipe):28s} {d:4d} {len(run['iterations']):4d}   -- "
                             f"oracle raised {type(e).__name__}: {e}")
            print(lines[-1], flush=True)
    store["files"] = np.array(names)
    np.savez_compressed(args.out, **store)
    with open(args.report, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(f"wrote {args.out} ({os.path.getsize(args.out) / 1024:.0f} KiB) and {args.report}")


if __name__ == "__main__":
    main()
