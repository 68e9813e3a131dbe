#!/usr/bin/env python3
"""TK_TTR_ONEROUND=1 (experimental one-reduction-round 3-term step, tk_krylov.cuh) against the default two-round
kernel: Lanczos coefficients, b~ and basis vectors on a few operators, then the 3-term kernel's time inside a C5
solve with and without it.  Needs a GPU.  Writes gpurun_out/oneround.txt line by line."""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry   # noqa: E402

tk = entry.load_package()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "oneround.txt"), "w")


def say(msg):
    print(msg)
    out.write(msg + "\n")
    out.flush()


def bases(A, b, nmax, oneround, cpm):
    os.environ["TK_TTR_ONEROUND"] = "1" if oneround else "0"
    os.environ["TK_TTR_CPM"] = str(cpm)
    d = len(b)
    s = tk.Solver(d, A.shape[0], nmax, tk.SymInstance, tk.RandSPD, tk.TensorLanczos)
    s.set_operators([A] * d)
    s.set_rhs(b)
    s.begin()
    for k in range(2, nmax + 1):
        s.step_bases(k)
    res = [(s.get_H(m), s.get_bt(m), s.get_V(m, nmax // 2), s.get_V(m, nmax + 1)) for m in range(d)]
    s.close()
    return res


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def main():
    rng = np.random.default_rng(7)
    t0 = time.time()
    for name, n, cpm in [("laplace", 1000, 1), ("laplace", 1001, 2), ("laplace", 4000, 4), ("varying3", 2000, 2),
                         ("banded4", 1500, 2), ("clustered", 1200, 1)]:
        if name == "laplace":
            A = tk.assemble_matrix(n, tk.Laplace)
        elif name == "varying3":      # non-constant tridiagonal: the kernel reads the diagonals
            dg = 2.0 + rng.random(n); off = -1.0 + 0.3 * rng.random(n - 1)
            A = sp.diags([off, dg, off], [-1, 0, 1]).tocsc()
        elif name == "banded4":       # 4 diagonals (ND = 4), symmetric part only matters for the recurrence to be defined
            A = sp.diags([-np.ones(n - 1), 3.0 * np.ones(n), -np.ones(n - 1), 0.25 * np.ones(n - 2)], [-1, 0, 1, 2]).tocsc()
        else:                         # tridiagonal with a huge diagonal spread: beta << alpha, exercises the exact fallback
            dg = np.concatenate([np.full(n // 2, 1e6), np.full(n - n // 2, 1.0)]) + rng.random(n)
            A = sp.diags([-1e-3 * np.ones(n - 1), dg, -1e-3 * np.ones(n - 1)], [-1, 0, 1]).tocsc()
        d, nmax = 3, 30
        b = [v / np.linalg.norm(v) for v in (rng.random(n) for _ in range(d))]
        one, two = bases(A, b, nmax, True, cpm), bases(A, b, nmax, False, cpm)
        eH = max(rel(o[0], t[0]) for o, t in zip(one, two))
        eb = max(rel(o[1], t[1]) for o, t in zip(one, two))
        eV = max(max(np.max(np.abs(o[2] - t[2])), np.max(np.abs(o[3] - t[3]))) for o, t in zip(one, two))
        say(f"{name:10s} n={n:5d} cpm={cpm}: rel dev H {eH:.2e}  b~ {eb:.2e}  max |dV| {eV:.2e}")
    say(f"correctness part: {time.time() - t0:.1f} s")
    # timing inside a C5 solve (d = 1024, n = 10^4, nmax = 64, TensorLanczosReorth, fixed iterations)
    os.environ.pop("TK_TTR_CPM", None)
    d, n, nmax = 1024, 10000, 64
    A = tk.assemble_matrix(n, tk.Laplace)
    bvec = np.random.default_rng(12345).random(n)
    bvec /= np.linalg.norm(bvec)
    for oneround in (0, 1, 0, 1):
        os.environ["TK_TTR_ONEROUND"] = str(oneround)
        s = tk.Solver(d, n, nmax, tk.SymInstance, tk.Laplace, tk.TensorLanczosReorth,
                      flags=tk.TK_FLAG_REFERENCE_H1 | tk.TK_FLAG_FIXED_ITERATIONS | tk.TK_FLAG_TIME_KERNELS)
        s.set_operators([A] * d)
        s.set_rhs([bvec] * d)
        s.set_schedule(A, 1e-8)
        s.solve(1e-8)
        res = s.solve(1e-8)
        ms_ttr, nl, by = s.timing(0)
        ms_gram, _, byg = s.timing(1)
        ms_all, _, _ = s.timing(6)
        say(f"oneround={oneround}: solve {ms_all:.2f} ms; 3-term {ms_ttr:.3f} ms / {nl} launches = {by / ms_ttr / 1e6:.0f} GB/s; "
            f"gram {ms_gram:.2f} ms = {byg / ms_gram / 1e6:.0f} GB/s; relres[-1] {res['relres'][-1]:.6e}")
        s.close()
    say(f"total {time.time() - t0:.1f} s")


if __name__ == "__main__":
    main()
