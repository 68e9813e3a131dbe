#!/usr/bin/env python3
"""Per-phase device time of one solve (TK_FLAG_TIME_ALL): python tools/profile_phases.py [d n nmax variant per_mode]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as entry
tk = entry.load_package()
d = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
nmax = int(sys.argv[3]) if len(sys.argv) > 3 else 64
vname = sys.argv[4] if len(sys.argv) > 4 else "reorth"
variant = {"reorth": tk.TensorLanczosReorth, "lanczos": tk.TensorLanczos, "arnoldi": tk.TensorArnoldi}[vname]
inst, cls = (tk.NonSymInstance, tk.ConvDiff) if vname == "arnoldi" else (tk.SymInstance, tk.Laplace)
per_mode = len(sys.argv) > 5 and sys.argv[5] == "1"
flags = tk.TK_FLAG_FIXED_ITERATIONS | tk.TK_FLAG_TIME_ALL | tk.TK_FLAG_TIME_KERNELS | (0 if per_mode else tk.TK_FLAG_REFERENCE_H1)
A1 = tk.assemble_matrix(n, cls)
b = np.random.default_rng(12345).random(n); b /= np.linalg.norm(b)
s = tk.Solver(d, n, nmax, inst, cls, variant, flags=flags)
s.set_operators([A1] * d); s.set_rhs([b] * d); s.set_schedule(A1, 1e-8)
for _ in range(3):
    r = s.solve(1e-8)
names = ["ttr", "gram", "mgs", "eig|expm", "assemble+gramblocks", "combine+finalize", "solve"]
out = {}
for i, nm in enumerate(names):
    ms, cnt, by = s.timing(i)
    out[nm] = {"ms": round(ms, 3), "launches": cnt, "GB/s": round(by / ms / 1e6, 1) if ms > 0 and by > 0 else None}
out["launch_count"] = s.launch_count()
out["relres_last"] = float(r["relres"][nmax - 1])
print(json.dumps({"d": d, "n": n, "nmax": nmax, "per_mode": per_mode, **out}))
