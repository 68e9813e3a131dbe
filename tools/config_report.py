#!/usr/bin/env python3
"""Runs BASELINE.json's five configurations through the public API on one B200 and, where the CPU oracle
finishes in reasonable time, times the oracle beside it.  Writes one JSON object per configuration."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as entry
tk = entry.load_package()
orc = entry.load_oracle()
tables = orc.ExpSumTables.from_packed(tk.TABLES_PATH)
CPU = "--cpu" in sys.argv

def run(name, d, n, nmax, tol, cls, inst, variant, ocls, oinst, ovariant, cpu_mode=None, cpu_iters=None):
    b = np.random.default_rng(12345).random(n)
    A = tk.KroneckerMatrix.gallery(inst, d, n, cls)
    system = tk.TensorizedSystem(inst, A, [b] * d)
    out = []
    best = None
    for rep in range(3):
        t0 = time.perf_counter()
        cd = tk.ConvergenceData(nmax)
        slv = []
        tk.tensorkrylov(cd, system.A, system.b, tol, nmax, variant, verbose=False, solver_out=slv)
        wall = time.perf_counter() - t0
        dev_ms = slv[0].timing(6)[0]
        slv[0].close()
        best = wall if best is None else min(best, wall)
    rr = cd.relative_residual_norm
    rec = dict(config=name, d=d, n=n, nmax=nmax, tol=tol, status=cd.status, term_k=cd.term_k, niterations=cd.niterations,
               best_relres=float(rr[1:].min()) if len(rr) > 1 else None, best_k=int(rr[1:].argmin()) + 2 if len(rr) > 1 else None,
               gpu_wall_ms=1e3 * best, gpu_device_ms=dev_ms,
               gpu_iters_per_s=(cd.term_k - 1) / (dev_ms / 1e3) if cd.term_k and cd.term_k > 1 else None)
    if CPU and cpu_mode:
        Ao = orc.assemble_matrix(n, ocls)
        kw = dict(residual="faithful") if cpu_mode == "A" else dict(residual="nilpotent", fast_solve=(oinst == orc.SYM))
        t0 = time.perf_counter()
        S = orc.OracleSolve([Ao] * d, orc.normalize_rhs([b] * d), tol, nmax, ovariant, oinst, ocls, tables, **kw)
        nit = 0
        while S.status is None and (cpu_iters is None or nit < cpu_iters):
            S.iterate(); nit += 1
        dt = time.perf_counter() - t0
        rec.update(cpu_flavour=cpu_mode, cpu_iterations_timed=nit, cpu_s=dt, cpu_iters_per_s=nit / dt, cpu_threads=os.cpu_count(),
                   cpu_status=S.status, cpu_term_k=S.k)
        kk = np.arange(2, min(S.k, cd.term_k or nmax, 40) + 1)
        rec["max_rel_dev_relres_k<=40"] = float(np.max(np.abs(rr[kk - 1] - S.relres[kk - 1]) / S.relres[kk - 1]))
    print(json.dumps(rec), flush=True)

L, R, Ar = tk.TensorLanczos, tk.TensorLanczosReorth, tk.TensorArnoldi
run("C1", 5, 200, 199, 1e-8, tk.Laplace, tk.SymInstance, R, orc.LAPLACE, orc.SYM, orc.LANCZOS_REORTH, "A")
run("C2", 50, 1000, 256, 1e-8, tk.Laplace, tk.SymInstance, R, orc.LAPLACE, orc.SYM, orc.LANCZOS_REORTH, "B", 64)
run("C3-1gpu", 256, 10000, 64, 1e-8, tk.Laplace, tk.SymInstance, R, orc.LAPLACE, orc.SYM, orc.LANCZOS_REORTH, None)
run("C3-1gpu-tol1e-5", 256, 10000, 64, 1e-5, tk.Laplace, tk.SymInstance, R, orc.LAPLACE, orc.SYM, orc.LANCZOS_REORTH, None)
run("C4", 100, 2000, 120, 1e-8, tk.ConvDiff, tk.NonSymInstance, Ar, orc.CONVDIFF, orc.NONSYM, orc.ARNOLDI, "B", 24)
run("C5-tol1e-5", 1024, 10000, 64, 1e-5, tk.Laplace, tk.SymInstance, R, orc.LAPLACE, orc.SYM, orc.LANCZOS_REORTH, None)
run("C5", 1024, 10000, 64, 1e-8, tk.Laplace, tk.SymInstance, R, orc.LAPLACE, orc.SYM, orc.LANCZOS_REORTH, None)
