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
