import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
tk = entry.load_package()
import ctypes as C
lib = tk._capi.lib
d, n, nmax = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024), 10000, 64
A1 = tk.assemble_matrix(n, tk.Laplace)
b = np.random.default_rng(12345).random(n); b /= np.linalg.norm(b)
flags = tk.TK_FLAG_FIXED_ITERATIONS | tk.TK_FLAG_REFERENCE_H1
for it in range(4):
    t = [time.perf_counter()]
    s = tk.Solver(d, n, nmax, tk.SymInstance, tk.Laplace, tk.TensorLanczosReorth, flags=flags); t.append(time.perf_counter())
    s.set_operators([A1] * d); t.append(time.perf_counter())
    s.set_rhs([b] * d); t.append(time.perf_counter())
    s.set_schedule(A1, 1e-8); t.append(time.perf_counter())
    r = s.solve(1e-8); t.append(time.perf_counter())
    dev = s.timing(6)[0]
    s.close(); t.append(time.perf_counter())
    names = ["create", "set_operators", "set_rhs", "set_schedule", "solve", "close"]
    print(it, {nm: round(1e3 * (t[i + 1] - t[i]), 2) for i, nm in enumerate(names)}, "device solve ms", round(dev, 2))
