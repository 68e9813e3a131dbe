import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as entry
tk = entry.load_package()
d = int(sys.argv[1]); n = 10000; nmax = 64
A1 = tk.assemble_matrix(n, tk.Laplace)
b = np.random.default_rng(12345).random(n); b /= np.linalg.norm(b)
s = tk.Solver(d, n, nmax, tk.SymInstance, tk.Laplace, tk.TensorLanczosReorth, flags=tk.TK_FLAG_REFERENCE_H1)
s.set_operators([A1] * d); s.set_rhs([b] * d); s.set_schedule(A1, 1e-5)
for i in range(int(sys.argv[2])):
    r = s.solve(1e-5)
print("ok", d, r["status"], r["term_k"], float(r["relres"][r["term_k"] - 1]))
