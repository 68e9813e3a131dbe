"""Host mirror of the reference's experiment drivers (tensorkrylov.jl_b200/experiments.py): the operator families equal
the recipes that reproduce the stored Julia runs (tools/corpus_sweep.py, validated by tests/test_oracle_corpus.py), the
decoder of Julia-serialized `Experiment` objects reads every stored file, and the npz round trip keeps everything."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import corpus as C

REF = os.environ.get("TK_REFERENCE", "/root/reference")
DATA = os.path.join(REF, "experiments", "data")


def dense(M):
    return M.toarray() if sp.issparse(M) else np.asarray(M)


def test_operator_families_equal_the_validated_recipes(tk, orc):
    ex, cs = tk.experiments, C.corpus_sweep
    for alpha in (1.9999, 2.005, 1.999756):
        assert np.array_equal(dense(ex.parameterize(alpha, tk.SymInstance)), dense(cs.sym_parameterized(alpha)))
    for beta in (-3.0, -5.005, -5.07):
        assert np.array_equal(dense(ex.parameterize(beta, tk.NonSymInstance)), dense(cs.nonsym_parameterized(beta)))
    # beta = -5 is the gallery's ConvDiff (tensor_struct.jl:60-68), alpha = 2 its Laplace
    assert np.array_equal(dense(ex.parameterize(-5.0, tk.NonSymInstance)), dense(tk.assemble_matrix(200, tk.ConvDiff)))
    assert np.array_equal(dense(ex.parameterize(2.0, tk.SymInstance)), dense(tk.assemble_matrix(200, tk.Laplace)))
    assert np.array_equal(ex.clusterzero(200), cs.clusterzero()) and np.array_equal(ex.clusterone(200), cs.clusterone())
    for d in (5, 10):
        A = ex.uniform_kroneckersum(200, d, (1e-3, 1.0))
        ref, _ = cs.operators(orc, ("eig_uniform", 1e-3, 1.0), d)
        assert len(A) == d and all(np.array_equal(A[s], ref[s]) for s in range(d))
        assert A.matrixclass is tk.EigValMat and A.instance is tk.SymInstance
        P = tk.KroneckerMatrix(tk.SymInstance, [tk.assemble_matrix(ex.clusterone(200), tk.EigValMat)] * d, tk.EigValMat)
        ex.perturb_matrix(P, 1e-2)
        ref, _ = cs.operators(orc, ("eig_one_eps", 1e-2), d)
        assert all(np.array_equal(P[s], ref[s]) for s in range(d))
        assert not np.array_equal(P[0], P[1])          # a different dense matrix in every mode


def test_experiment_record_and_npz_round_trip(tk, tmp_path):
    ex = tk.experiments
    rng = np.random.default_rng(3)
    b = ex.multiple_rhs([5, 10], 200, rng)
    assert [len(x) for x in b] == [5, 10] and all(v is b[1][0] for v in b[1])      # one vector, d times (system.jl:5-11)
    e = ex.Experiment([5, 10], 200, 30, tk.SymInstance, tk.Laplace, tk.TensorLanczosReorth, b)
    assert len(e) == 2 and [c.niterations for c in e.conv_vector] == [30, 30]
    assert np.all(ex.get_relative_residuals(e)[0] == 1.0) and ex.get_max_iteration(e) == 30
    e.conv_vector[1].relative_residual_norm[3] = 0.25
    path = tmp_path / "exp.npz"
    ex.serialize_to_file(path, e)
    z = np.load(path)
    assert list(z["dims"]) == [5, 10] and list(z["tags"]) == ["SymInstance", "Laplace", "TensorLanczosReorth"]
    assert np.array_equal(z["rhs_d10"], b[1][0]) and z["relres_d10"][3] == 0.25 and int(z["niterations_d5"]) == 30


@pytest.mark.skipif(not os.path.isdir(DATA), reason="reference tree not present")
def test_decoder_reads_every_stored_file(tk):
    """All 30 files under experiments/data: type tags, 5+10+50+100 right-hand sides, four ConvergenceData records; the
    27 reproducible ones equal tests/golden/corpus.npz bit for bit."""
    z = C.corpus()
    seen = 0
    for sub in sorted(os.listdir(DATA)):
        for name in sorted(os.listdir(os.path.join(DATA, sub))):
            e = tk.experiments.deserialize_from_file(os.path.join(DATA, sub, name))
            seen += 1
            assert e.dims == [5, 10, 50, 100] and e.matrixsize == 200
            assert [len(r) for r in e.rhs_vec] == e.dims
            for c in e.conv_vector:
                assert len(c.relative_residual_norm) == len(c.iterations) == c.niterations
                assert c.relative_residual_norm[0] == 1.0 and list(c.iterations[:2]) == [1, 2][: len(c.iterations)]
            key = f"{sub}__{name}"
            if key in C.files():
                assert [e.instance.__name__, e.matrixclass.__name__, e.orth_method.__name__] == [str(x) for x in z[f"{key}__meta"][:3]]
                for d, rhs, c in zip(e.dims, e.rhs_vec, e.conv_vector):
                    assert np.array_equal(rhs[0], z[f"{key}__rhs_d{d}"])
                    assert np.array_equal(c.relative_residual_norm, z[f"{key}__relres_d{d}"])
                    assert np.array_equal(c.projected_residual_norm, z[f"{key}__projres_d{d}"])
                    assert np.array_equal(c.orthogonality_data, z[f"{key}__orth_d{d}"])
    assert seen == 30


def test_driver_plumbing_with_an_oracle_backed_solver(tk, orc, tables, monkeypatch):
    """The checks of the GPU driver test (tests/test_zz_gpu_corpus.py), with `solve_tensorized_system` -- the one call
    that needs the GPU, and which the GPU suite tests on its own -- replaced by the CPU oracle: exercises everything the
    drivers do around it (operators, class tags, right-hand sides, keyword plumbing, ConvergenceData bookkeeping)."""
    cls_of = {tk.Laplace: orc.LAPLACE, tk.ConvDiff: orc.CONVDIFF, tk.RandSPD: orc.RANDSPD, tk.EigValMat: orc.EIGVALMAT}
    var_of = {tk.TensorLanczos: orc.LANCZOS, tk.TensorLanczosReorth: orc.LANCZOS_REORTH, tk.TensorArnoldi: orc.ARNOLDI}
    calls = []

    def stand_in(system, nmax, orth, tol=1e-9, verbose=True, **kw):
        assert not kw and verbose is False
        inst = orc.NONSYM if system.instance is tk.NonSymInstance else orc.SYM
        cls = cls_of[system.A.matrixclass]
        A = [sp.csr_matrix(M) if sp.issparse(M) else np.asarray(M) for M in system.A.M]
        assert all(abs(np.linalg.norm(b) - 1.0) < 1e-15 for b in system.b)          # TensorizedSystem normalised them
        S = orc.tensorkrylov(A, system.b, tol, nmax, var_of[orth], inst, cls, tables if inst == orc.SYM else None,
                             fast_solve=(cls != orc.EIGVALMAT and inst == orc.SYM))
        cd = tk.ConvergenceData(nmax)
        cd.status, cd.niterations = S.status, S.niterations
        cd.relative_residual_norm[:len(S.relres)] = S.relres
        calls.append((system.d, system.A.matrixclass.__name__, orth.__name__))
        return cd

    monkeypatch.setattr(tk.experiments, "solve_tensorized_system", stand_in)
    C.check_experiment_drivers(tk, tol=1e-11)
    assert len(calls) == 2 * (4 + 4 + 4) + 2 and {c[0] for c in calls} == {5, 10}
    assert {c[1:] for c in calls} == {("Laplace", "TensorLanczosReorth"), ("ConvDiff", "TensorArnoldi"),
                                      ("RandSPD", "TensorLanczosReorth"), ("EigValMat", "TensorLanczosReorth")}
