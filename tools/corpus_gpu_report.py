#!/usr/bin/env python3
"""CUDA path vs the reference's stored Julia histories over the whole regression corpus (tests/golden/corpus.npz:
27 files x d = 5, 10, 50, 100), through the C-ABI.  Same quantities as tools/corpus_sweep.py reports for the CPU
oracle.  Needs a GPU; reads nothing from the reference.  Usage: python tools/corpus_gpu_report.py [--kmax 40] [--out F]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry   # noqa: E402
import corpus as C                # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kmax", type=int, default=40)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "corpus_gpu_report.txt"))
    ap.add_argument("--budget", type=float, default=1e9, help="stop after this many seconds (partial report)")
    args = ap.parse_args()
    tk, orc = entry.load_package(), entry.load_oracle()
    t0 = time.time()
    lines = ["# CUDA path (C-ABI, TK_FLAG_REFERENCE_H1 | TK_FLAG_FIXED_ITERATIONS) vs the reference's stored Julia histories",
             "# dev2: max_k |relres^2 - stored^2| (||b|| = 1);  e_rel: max_k |relres - stored|/stored over iterations with stored^2 > 1e-3",
             f"{'file':36s} {'d':>4s} {'K':>3s} {'dev2':>10s} {'e_rel':>10s} {'ms':>8s}"]
    cls_of = {"Laplace": tk.Laplace, "ConvDiff": tk.ConvDiff, "RandSPD": tk.RandSPD, "EigValMat": tk.EigValMat}
    var_of = {"TensorLanczos": tk.TensorLanczos, "TensorLanczosReorth": tk.TensorLanczosReorth, "TensorArnoldi": tk.TensorArnoldi}
    for d in (5, 10, 50, 100):
        for key in C.files():
            if time.time() - t0 > args.budget:
                break
            e = C.entry(key, d)
            K = min(args.kmax, e["length"])
            if K < 2:
                continue
            A, _ = C.corpus_sweep.operators(orc, e["recipe"], d)
            inst = tk.NonSymInstance if e["instance"] == "NonSymInstance" else tk.SymInstance
            b = e["rhs"] * (1.0 / np.linalg.norm(e["rhs"]))
            t1 = time.time()
            slv = tk.Solver(d, 200, K, inst, cls_of[e["cls"]], var_of[e["orth"]],
                            flags=tk.TK_FLAG_REFERENCE_H1 | tk.TK_FLAG_FIXED_ITERATIONS)
            try:
                slv.set_operators(A)
                slv.set_rhs([b] * d)
                slv.set_schedule(A[0], 1e-9)
                res = slv.solve(1e-9)
            finally:
                slv.close()
            ms = (time.time() - t1) * 1e3
            k = np.arange(2, K + 1)
            got, ref = res["relres"][k - 1], e["relres"][k - 1]
            dev2 = np.abs(got ** 2 - ref ** 2).max()
            well = ref ** 2 > 1e-3
            e_rel = (np.abs(got - ref)[well] / ref[well]).max() if well.any() else float("nan")
            lines.append(f"{key:36s} {d:4d} {K:3d} {dev2:10.2e} {e_rel:10.2e} {ms:8.1f}")
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(f"{len(lines) - 3} runs in {time.time() - t0:.1f} s -> {args.out}")


if __name__ == "__main__":
    main()
