"""Shared by the CPU and GPU corpus tests: tests/golden/corpus.npz holds, for every reproducible file under the
reference's experiments/data (27 of 30), the right-hand side the authors drew and the histories their Julia run
stored, for d = 5, 10, 50, 100 at n = 200 (tools/corpus_sweep.py wrote it and documents the decoding and how the
drivers' unstored parameters were recovered)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import corpus_sweep  # noqa: E402  (operator recipes of the experiment drivers; reads nothing at import)

_Z = None


def corpus():
    global _Z
    if _Z is None:
        _Z = np.load(os.path.join(ROOT, "tests", "golden", "corpus.npz"))
    return _Z


def files():
    return [str(f) for f in corpus()["files"]]


def entry(key, d):
    """-> dict(instance, cls, orth, recipe, rhs, relres, projres, orth_hist, length) for one stored run."""
    z = corpus()
    inst, cls, orth, kind = (str(x) for x in z[f"{key}__meta"])
    recipe = (kind,) + tuple(float(x) for x in z[f"{key}__param"])
    return dict(instance=inst, cls=cls, orth=orth, recipe=recipe, rhs=z[f"{key}__rhs_d{d}"],
                relres=z[f"{key}__relres_d{d}"], projres=z[f"{key}__projres_d{d}"], orth_hist=z[f"{key}__orth_d{d}"],
                length=len(z[f"{key}__iterations_d{d}"]))


def check_experiment_drivers(tk, dims=(5, 10), K=30, tol=4e-11):
    """The reference-shaped experiment entry points on the stored right-hand sides against the stored Julia histories.
    Every solve goes through `tk.experiments.solve_tensorized_system` (the library on the GPU box; the CPU test swaps in
    an oracle-backed stand-in to exercise the driver plumbing)."""
    ex = tk.experiments

    def rhs(key):
        return [[entry(key, d)["rhs"]] * d for d in dims]

    def close(cd, key, d):
        k = np.arange(2, K + 1)
        ref = entry(key, d)["relres"]
        assert cd.status == tk.TK_NMAX and cd.niterations == K, (key, d, cd.status)      # no breakdown this early
        dev = np.abs(cd.relative_residual_norm[k - 1] ** 2 - ref[k - 1] ** 2).max()
        assert dev <= tol, (key, d, dev)

    spd, _ = ex.reproduce(200, 1e-9, dims, nmax=K, rhs=rhs("reproduction_data__laplace_new"), verbose=False)
    _, nonsym = ex.reproduce(200, 1e-9, dims, nmax=K, rhs=rhs("reproduction_data__nonsym_new"), verbose=False)
    for i, d in enumerate(dims):
        close(spd.conv_vector[i], "reproduction_data__laplace_new", d)
        close(nonsym.conv_vector[i], "reproduction_data__nonsym_new", d)
    spd, _ = ex.parameterized_experiment(1.99976, -5.05, 1e-9, dims, nmax=K, rhs=rhs("parametrized_data__sym4"), verbose=False)
    _, nonsym = ex.parameterized_experiment(1.99976, -5.05, 1e-9, dims, nmax=K, rhs=rhs("parametrized_data__nonsym4"),
                                            verbose=False)
    for i, d in enumerate(dims):
        close(spd.conv_vector[i], "parametrized_data__sym4", d)
        close(nonsym.conv_vector[i], "parametrized_data__nonsym4", d)
    zero, _ = ex.eigenvalue_experiment(200, rhs("eigenvalues_data__d2zero"), 1e-2, 1e-9, dims, nmax=K, perturb=True, verbose=False)
    _, one = ex.eigenvalue_experiment(200, rhs("eigenvalues_data__d2one"), 1e-2, 1e-9, dims, nmax=K, perturb=True, verbose=False)
    uni = ex.uniform_experiment(dims, 200, rhs("eigenvalues_data__uniform"), (1e-3, 1.0), 1e-9, nmax=K, verbose=False)
    for i, d in enumerate(dims):
        close(zero.experiment.conv_vector[i], "eigenvalues_data__d2zero", d)
        close(one.experiment.conv_vector[i], "eigenvalues_data__d2one", d)
        close(uni.experiment.conv_vector[i], "eigenvalues_data__uniform", d)
