#!/usr/bin/env python3
"""How much the MGS fallback of TensorLanczosReorth costs when it fires often.

Operators with the clustered spectrum j^2/n^2 (the reference's eigenvalues_data runs, scaled up): the
orthogonality monitor trips from k ~ 60 on and the fallback then runs at almost every step, as a CTA-wide two-pass
MGS inside the Gram-row kernel (monitor_body -> mgs_step_cta, one 256-thread CTA per mode).  Prints, per size, the
solve time, the time in the Gram-row kernel (which includes the fallback) and the number of fallbacks of mode 0.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    tk = entry.load_package()
    out = []
    import scipy.sparse as sp
    for d, n, nmax in ((5, 200, 199), (128, 2000, 200), (128, 10000, 200)):
        # the clustered spectrum as a SPARSE diagonal operator under the RandSPD spectral rule: cheap SpMV and the
        # symmetric compressed solve, so the Krylov-step kernels (with the fallback inside) are what is timed
        ev = (np.arange(1, n + 1) / float(n)) ** 2
        A = sp.diags(ev).tocsc()
        b = np.random.default_rng(12345).random(n)
        b /= np.linalg.norm(b)
        row = {"d": d, "n": n, "nmax": nmax}
        for variant, name in ((tk.TensorLanczosReorth, "reorth"), (tk.TensorLanczos, "lanczos")):
            s = tk.Solver(d, n, nmax, tk.SymInstance, tk.RandSPD, variant,
                          flags=tk.TK_FLAG_FIXED_ITERATIONS | tk.TK_FLAG_TIME_KERNELS | tk.TK_FLAG_REFERENCE_H1)
            s.set_operators([A] * d)
            s.set_rhs([b] * d)
            s.set_schedule(A, 1e-9)
            s.solve(1e-9)
            s.solve(1e-9)
            row[name] = {"solve_ms": s.timing(6)[0], "gram_row_ms": s.timing(1)[0], "three_term_ms": s.timing(0)[0],
                         "mgs_fallbacks_mode0": s.orth_state(0)[1]}
            s.close()
        out.append(row)
        print(json.dumps(row), flush=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_fallback_report.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
