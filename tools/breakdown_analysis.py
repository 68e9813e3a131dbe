#!/usr/bin/env python3
"""Where the reference's stored runs ended with `CompressedNormBreakdown` (r_comp < 0, utils.jl:395): what the CPU
oracle computes at that very iteration, next to the last values the Julia run stored.

Reads tests/golden/corpus.npz only.  Output: profiles/r01_breakdown_analysis.txt.  The point it documents: at the
iteration where Julia's r_comp turned negative the quantity is a cancellation residue at the level of Julia's own
summation noise; the oracle (and the CUDA path, which shares its O(d t^2) combine) still has r_comp > 0 there, so
"the same breakdown iteration" is not a property any other summation order can reproduce."""
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
    orc = entry.load_oracle()
    tk_tables = os.path.join(ROOT, "tensorkrylov.jl_b200", "data", "expsum_tables.bin")
    tables = orc.ExpSumTables.from_packed(tk_tables)
    cs = C.corpus_sweep
    budget = float(os.environ.get("TK_BUDGET_S", "1e9"))
    lines = ["# stored Julia runs that ended in CompressedNormBreakdown: L = stored history length, Julia's r_comp < 0 at k = L+1",
             "# julia[L-4..L]: min / max of the last five stored r_comp;  oracle(L+1): the oracle's r_comp at Julia's breakdown iteration",
             "# max |diff|: max |oracle - julia| over those five iterations -- as large as r_comp itself: by then r_comp is a cancellation residue of O(1) terms in both",
             f"{'file':34s} {'d':>4s} {'L':>4s} {'julia min':>10s} {'julia max':>10s} {'oracle(L+1)':>12s} {'max |diff|':>10s}"]
    t0 = time.time()
    for key in C.files():
        for d in (5, 10, 50, 100):
            e = C.entry(key, d)
            L = e["length"]
            if L >= 199 or L < 6 or time.time() - t0 > budget:
                continue
            if e["cls"] == "EigValMat" and d >= 50:
                continue                      # per-mode dense exponentials at d >= 50: minutes each on the CPU
            A, cls = cs.operators(orc, e["recipe"], d)
            inst = orc.NONSYM if e["instance"] == "NonSymInstance" else orc.SYM
            b = orc.normalize_rhs([e["rhs"]] * d)
            S = orc.OracleSolve(A, b, 1e-9, L + 1, cs.VARIANT[e["orth"]], inst, cls, tables if inst == orc.SYM else None,
                                ignore_breakdown=True, fast_solve=(cls != orc.EIGVALMAT and inst == orc.SYM))
            S.run()
            k = np.arange(L - 4, L + 1)
            st = e["projres"][k - 1]
            oc = np.array([S.detail[int(kk)]["r_comp"] for kk in k])
            lines.append(f"{key:34s} {d:4d} {L:4d} {st.min():10.2e} {st.max():10.2e} {S.detail[L + 1]['r_comp']:12.2e} "
                         f"{np.abs(st - oc).max():10.2e}")
            print(lines[-1], flush=True)
    with open(os.path.join(ROOT, "profiles", "r01_breakdown_analysis.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
