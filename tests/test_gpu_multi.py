"""Multi-GPU parity (needs >= 2 B200s on the box; skipped otherwise): the mode-sharded NCCL solve reproduces the
single-GPU solve -- histories, status and every rank's local solution factors."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
import __graft_entry__ as entry
import ctypes as C
tk = entry.load_package()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    raw = C.create_string_buffer(128)
    tk._capi.check(tk._capi.lib.tk_comm_unique_id(raw))
    buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
dist.broadcast(buf, 0)
uid = bytes(buf.cpu().numpy().tobytes())
d, n, nmax, tol = {d}, {n}, {nmax}, {tol}
rng = np.random.default_rng(5)
NONSYM = {nonsym}
inst, cls, variant = (tk.NonSymInstance, tk.ConvDiff, tk.TensorArnoldi) if NONSYM else (tk.SymInstance, tk.Laplace, tk.TensorLanczosReorth)
A1 = tk.assemble_matrix(n, cls)
if {distinct}:
    b = [v / np.linalg.norm(v) for v in (rng.random(n) for _ in range(d))]
else:
    one = rng.random(n); one /= np.linalg.norm(one)
    b = [one] * d
flags = {flags}
out = {{}}
for label, w, r, u in (("multi", world, rank, uid), ("multi_nccl", world, rank, uid), ("single", 1, 0, None)):
    if label == "single" and rank != 0:
        continue
    os.environ["TK_PEER"] = "0" if label == "multi_nccl" else "1"     # read when the exchange buffers are set up
    s = tk.Solver(d, n, nmax, inst, cls, variant, flags=flags, device=rank, rank=r, world=w,
                  unique_id=u)
    s.set_operators([A1] * d); s.set_rhs(b); s.set_schedule(A1, tol)
    res = s.solve(tol)
    info = s.solve_info()
    lam, fm = s.solution(force=True)
    if label == "multi":
        # the second solve of the handle replays CUDA graphs (with the peer exchange inside): identical histories
        res2 = s.solve(tol)
        info2 = s.solve_info()
        assert info2["graphs_launched"] > 0, info2
        assert np.array_equal(res2["relres"], res["relres"]) and res2["term_k"] == res["term_k"]
    out[label] = (res, lam, fm, s.first, s.count, info)
    s.close()
# the exchange through peer-mapped memory and the NCCL all-gather deliver the same partials: bit-identical solves
assert out["multi_nccl"][5]["peer_exchange"] is False
assert np.array_equal(out["multi"][0]["relres"], out["multi_nccl"][0]["relres"])
assert np.array_equal(out["multi"][0]["projres"], out["multi_nccl"][0]["projres"])
print("rank", rank, "peer exchange:", out["multi"][5]["peer_exchange"])
if rank == 0:
    rm, rs = out["multi"][0], out["single"][0]
    assert rm["status"] == rs["status"] and rm["term_k"] == rs["term_k"], (rm["status"], rs["status"])
    k = rs["term_k"]
    scale = 4.0
    assert np.max(np.abs(rm["relres"][:k] ** 2 - rs["relres"][:k] ** 2)) <= 1e-11 * scale
    assert np.allclose(out["multi"][1], out["single"][1], rtol=1e-14)
    first, count = out["multi"][3], out["multi"][4]
    for s_ in range(first, first + count):
        assert np.max(np.abs(out["multi"][2][s_] - out["single"][2][s_])) < 1e-9
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
"""


def _ngpus(tk):
    return tk.device_count()


@pytest.mark.parametrize("d,n,nmax,tol,distinct,ref_h1,nonsym", [(48, 600, 14, 1e-8, False, True, False),
                                                               (37, 400, 10, 1e-8, True, False, False),
                                                               (64, 1000, 40, 1e-4, False, True, False),
                                                               (20, 300, 12, 1e-8, False, True, True),
                                                               (9, 200, 10, 1e-8, True, False, True)])
def test_two_gpu_solve_equals_single_gpu(tk, gpu, tmp_path, d, n, nmax, tol, distinct, ref_h1, nonsym):
    if _ngpus(tk) < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    flags = tk.TK_FLAG_REFERENCE_H1 if ref_h1 else 0
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, d=d, n=n, nmax=nmax, tol=tol, distinct=distinct, flags=flags, nonsym=nonsym))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o[-3000:]
        assert f"rank {r} ok" in o
