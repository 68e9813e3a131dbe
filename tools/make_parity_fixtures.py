#!/usr/bin/env python3
"""Oracle fixtures at BASELINE.json's sizes (tests/golden/c{2,3,4,5}_oracle.npz).

The reference stores Julia results only for n = 200, so at n = 10^3..10^4 the CPU oracle (oracle/tk_oracle.py,
itself pinned to the stored Julia runs) is the available pin.  Inputs are the bench's: one U(0,1) vector from
numpy default_rng(12345) for every mode, normalised (random_rhs + TensorizedSystem, system.jl:5-43).

  C2  d=50,   n=1000,  Laplace,  TensorLanczosReorth, nmax=256, tol 1e-8   whole solve, parity mode
  C4  d=100,  n=2000,  ConvDiff, TensorArnoldi,       nmax=120, tol 1e-8   whole solve, parity mode
  C3  d=256,  n=10^4,  Laplace,  TensorLanczosReorth, first 16 iterations, fixed-iteration mode (as bench.py runs it)
  C5  d=1024, n=10^4,  Laplace,  TensorLanczosReorth, first 16 iterations, fixed-iteration mode

Each file holds, per iteration k = 2..K: hy2 = ||Hy||^2, hyb = <Hy,b>, bb = ||b~||^2, boundary, r_comp, relres, t,
lambda_min, and mode 1's H (diag/sub/super or the dense Hessenberg) and b~.

Usage: python tools/make_parity_fixtures.py [c2 c3 c4 c5]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

CASES = {
    "c2": dict(d=50, n=1000, cls="LAPLACE", variant="LANCZOS_REORTH", instance="SYM", nmax=256, tol=1e-8, fixed=False),
    "c4": dict(d=100, n=2000, cls="CONVDIFF", variant="ARNOLDI", instance="NONSYM", nmax=120, tol=1e-8, fixed=False),
    "c3": dict(d=256, n=10000, cls="LAPLACE", variant="LANCZOS_REORTH", instance="SYM", nmax=16, tol=1e-8, fixed=True),
    "c5": dict(d=1024, n=10000, cls="LAPLACE", variant="LANCZOS_REORTH", instance="SYM", nmax=16, tol=1e-8, fixed=True),
}


def run_case(orc, tables, c, threads):
    A = orc.assemble_matrix(c["n"], getattr(orc, c["cls"]))
    b1 = np.random.default_rng(12345).random(c["n"])
    b = orc.normalize_rhs([b1] * c["d"])
    S = orc.OracleSolve([A] * c["d"], b, c["tol"], c["nmax"], getattr(orc, c["variant"]), getattr(orc, c["instance"]),
                        getattr(orc, c["cls"]), tables, ignore_breakdown=c["fixed"], mode_threads=threads)
    S.run()
    ks = sorted(S.detail)
    out = {f: np.array([S.detail[k][f] for k in ks]) for f in ("hy2", "hyb", "bb", "boundary", "r_comp", "r_norm")}
    out["k"] = np.array(ks)
    out["t"] = np.array([S.detail[k]["t"] for k in ks])
    out["lambda_min"] = np.array([S.detail[k]["lambda_min"] for k in ks])
    out["relres"] = np.asarray(S.relres)
    out["projres"] = np.asarray(S.projres)
    out["orth"] = np.asarray(S.orth)
    out["status"] = np.array(S.status)
    out["niterations"] = np.array(S.niterations)
    kk = ks[-1]
    out["H1"] = S.H[0][: kk + 1, : kk + 1].copy()
    out["bt1"] = S.bt[0][: kk + 1].copy()
    out["fallbacks"] = np.array(S.stats.get("fallbacks", 0))
    out["meta"] = np.array(repr(c))
    return out


def main():
    which = sys.argv[1:] or list(CASES)
    orc = entry.load_oracle()
    tables = orc.ExpSumTables.from_packed(os.path.join(entry.PKG_DIR, "data", "expsum_tables.bin"))
    threads = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass
    for name in which:
        t0 = time.perf_counter()
        out = run_case(orc, tables, CASES[name], threads)
        path = os.path.join(ROOT, "tests", "golden", f"{name}_oracle.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: status={out['status']} iterations={len(out['k'])} relres[-1]={out['relres'][-1]:.6e} "
              f"{time.perf_counter() - t0:.1f} s -> {path}", flush=True)


if __name__ == "__main__":
    main()
