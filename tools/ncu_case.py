#!/usr/bin/env python3
"""One short solve for ncu captures of the secondary kernels (no torch: plain ctypes through the C-ABI).

    python tools/ncu_case.py c5 [nmax]     d=1024, n=10^4 Laplace / TensorLanczosReorth, then the Kruskal solution
    python tools/ncu_case.py c4 [nmax]     d=100, n=2000 ConvDiff / TensorArnoldi (Hessenberg exponentials), then x
    python tools/ncu_case.py nonsym200     d=5, n=200 ConvDiff / Arnoldi, nmax=150: up to 111 exp-sum terms
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "c5"
    tk = entry.load_package()
    if case == "c5":
        d, n, nmax, cls, var, inst, tol = 1024, 10000, 24, tk.Laplace, tk.TensorLanczosReorth, tk.SymInstance, 1e-8
    elif case == "c4":
        d, n, nmax, cls, var, inst, tol = 100, 2000, 40, tk.ConvDiff, tk.TensorArnoldi, tk.NonSymInstance, 1e-8
    else:
        d, n, nmax, cls, var, inst, tol = 5, 200, 150, tk.ConvDiff, tk.TensorArnoldi, tk.NonSymInstance, 1e-9
    if len(sys.argv) > 2:
        nmax = int(sys.argv[2])
    b = np.random.default_rng(12345).random(n)
    b *= 1.0 / np.linalg.norm(b)
    A1 = tk.assemble_matrix(n, cls)
    s = tk.Solver(d, n, nmax, inst, cls, var, flags=tk.TK_FLAG_REFERENCE_H1 | tk.TK_FLAG_FIXED_ITERATIONS)
    s.set_operators([A1] * d)
    s.set_rhs([b] * d)
    s.set_schedule(A1, tol)
    r = s.solve(tol)
    lam, fm = s.solution(force=True)
    print(case, "status", r["status"], "relres", r["relres"][nmax - 1], "rank", len(lam), "launches", s.launch_count())
    s.close()


if __name__ == "__main__":
    main()
