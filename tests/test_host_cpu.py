"""CPU-side tests: the C-ABI surface, the table/schedule logic in the library against the oracle,
the host mirror of the Julia interface, and the multi-rank partition + merge logic under gloo."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_cabi_exports_every_declared_symbol(tk):
    """Every function include/tensorkrylov_b200.h declares is exported by the shared library."""
    hdr = open(os.path.join(ROOT, "include", "tensorkrylov_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(tk_[A-Za-z0-9_]+)\s*\(", hdr))
    assert declared, "header parse failed"
    assert declared == set(tk.EXPORTS)
    for name in declared:
        assert hasattr(tk._capi.lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", tk.LIB_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf"\bT {name}\b", out), f"{name} not exported"


def test_library_is_sm100a_and_fails_loudly_without_gpu(tk):
    out = subprocess.run(["cuobjdump", "-lelf", tk.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    if tk.device_count() == 0:
        with pytest.raises(tk.TKError):
            tk.Solver(2, 10, 4, tk.SymInstance, tk.Laplace, tk.TensorLanczos)


def test_tables_match_oracle(tk, orc, tables):
    """tk_tables_sym_lookup == the oracle's restatement of approximation.jl:65-84, incl. kappa rounded DOWN,
    missing rows skipped upward, and the CSV's column-11 quirk (ranks jump 10 -> 12)."""
    rng = np.random.default_rng(0)
    kappas = np.concatenate([10 ** rng.uniform(0.31, 13.5, 300), [2.0, 9.99, 10.0, 57.3, 899.9999999999999, 900.0, 1e4]])
    seen_ranks = set()
    for kappa in kappas:
        for tol in (1e-4, 1e-8, 1e-9, 1e-12):
            try:
                t0, om0, al0, dg0, od0 = tables.sym_lookup(kappa, tol)
            except (RuntimeError, KeyError):
                with pytest.raises(tk.TKError):
                    tk.sym_lookup(kappa, tol)
                continue
            t, om, al, dg, od = tk.sym_lookup(kappa, tol)
            assert (t, dg, od) == (t0, dg0, od0)
            assert np.array_equal(om, om0) and np.array_equal(al, al0)
            seen_ranks.add(t)
    assert 11 not in seen_ranks and {10, 12} <= seen_ranks
    # golden from the survey's decode of the reference run: kappa of Laplace k=2 -> R = 3e0, t = 6 at tol 1e-9
    lmin, lmax = orc.laplace_extremes(5, 200, 2)
    assert tk.sym_lookup(lmax / lmin, 1e-9)[0::3] == (6, 3)


def test_nonsym_coefficients_match_oracle(tk, orc):
    for lmin, tol in [(2.046548e5, 1e-9), (661.19, 1e-9), (3.7e7, 1e-8), (1.0, 1e-3)]:
        r0, om0, al0 = orc.nonsym_coeffs(lmin, tol)
        r, om, al = tk.nonsym_coefficients(lmin, tol)
        assert r == r0 and len(om) == 2 * r + 1
        assert np.allclose(om, om0, rtol=1e-15, atol=0) and np.allclose(al, al0, rtol=1e-14, atol=1e-300)
    assert tk.nonsym_coefficients(2.046548e5, 1e-9)[0] == 19      # 39 terms at k=2 of nonsym_new d=5


def test_laplace_extremes_match_oracle(tk, orc):
    import ctypes as C
    for d, n, k in [(5, 200, 2), (1024, 10000, 64), (50, 1000, 256)]:
        a, b = C.c_double(), C.c_double()
        tk._capi.check(tk._capi.lib.tk_laplace_extremes(d, n, k, C.byref(a), C.byref(b)))
        lo, hi = orc.laplace_extremes(d, n, k)
        assert a.value == pytest.approx(lo, rel=1e-15) and b.value == pytest.approx(hi, rel=1e-15)


@pytest.mark.skipif(not os.path.isdir("/root/reference/coefficients_data"), reason="reference tree not present")
def test_packed_tables_equal_reference_directory(tk, orc, tables):
    """The packed file is the reference's coefficients_data/ parsed to Float64 -- both loaders agree."""
    raw = orc.ExpSumTables.from_reference_dir("/root/reference/coefficients_data")
    assert np.array_equal(raw.R, tables.R) and np.array_equal(raw.err, tables.err)
    assert raw.coeffs.keys() == tables.coeffs.keys()
    for key in list(raw.coeffs)[::97]:
        assert np.array_equal(raw.coeffs[key][0], tables.coeffs[key][0])
        assert np.array_equal(raw.coeffs[key][1], tables.coeffs[key][1])
    tk.load_tables("/root/reference/coefficients_data")      # the C++ loader of the raw directory
    try:
        t, om, al, dg, od = tk.sym_lookup(4321.0, 1e-9)
        t0, om0, al0, _, _ = tables.sym_lookup(4321.0, 1e-9)
        assert t == t0 and np.array_equal(om, om0) and np.array_equal(al, al0)
    finally:
        tk.load_tables()


def test_host_mirror_types(tk):
    n, d = 30, 3
    A = tk.KroneckerMatrix.gallery(tk.SymInstance, d, n, tk.Laplace)
    assert len(A) == d and A.dimensions() == [n] * d and A[0] is A[2]
    assert abs(A[0][0, 0] - 2 * (n + 1) ** 2) < 1e-9 and abs(A[0][1, 0] + (n + 1) ** 2) < 1e-9
    C = tk.assemble_matrix(n, tk.ConvDiff).toarray()
    h = 1.0 / (n + 1)
    assert C[0, 0] == pytest.approx(2 / h**2 + 3 * 10 / (4 * h)) and C[0, 2] == pytest.approx(10 / (4 * h))
    assert C[1, 0] == pytest.approx(-1 / h**2 + 10 / (4 * h)) and C[0, 1] == pytest.approx(-1 / h**2 - 50 / (4 * h))
    b = tk.random_rhs(d, n, np.random.default_rng(1))
    assert b[0] is b[1]
    sysm = tk.TensorizedSystem(tk.SymInstance, A, b)
    assert all(abs(np.linalg.norm(v) - 1) < 1e-15 for v in sysm.b) and sysm.b[0] is sysm.b[1]
    with pytest.raises(AssertionError):
        tk.TensorizedSystem(tk.SymInstance, A, b[:2])
    cd = tk.ConvergenceData(7)
    assert cd.niterations == 7 and list(cd.iterations) == list(range(1, 8)) and np.all(cd.relative_residual_norm == 1)
    cd.resize(3)
    assert len(cd.iterations) == len(cd.orthogonality_data) == 3
    x = tk.KruskalTensor([2.0], [np.array([[1.0], [2.0]]), np.array([[3.0], [5.0]])])
    assert np.array_equal(tk.kroneckervectorize(x), 2.0 * np.kron([3.0, 5.0], [1.0, 2.0]))


def test_partition_modes(tk):
    for d, w in [(1024, 8), (1024, 4), (256, 8), (50, 2), (5, 4), (100, 8), (17, 2)]:
        parts = tk.partition_modes(d, w)
        assert sum(c for _, c in parts) == d
        pos = 0
        for f, c in parts:
            assert f == min(pos, d) or c == 0
            pos += c
        if d >= w * 16:
            assert all(f % 16 == 0 for f, _ in parts)


GLOO_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
import __graft_entry__ as entry
orc = entry.load_oracle(); tk = entry.load_package()
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
d, n, nmax, tol = 40, 60, 8, 1e-8
tables = orc.ExpSumTables.from_packed(tk.TABLES_PATH)
rng = np.random.default_rng(3)
A = orc.assemble_matrix(n, orc.LAPLACE)
b = orc.normalize_rhs([rng.random(n) for _ in range(d)])
full = orc.tensorkrylov([A] * d, b, tol, nmax, orc.LANCZOS_REORTH, orc.SYM, orc.LAPLACE, tables, per_mode=True,
                        ignore_breakdown=True)
first, count = tk.partition_modes(d, 2)[rank]
# every rank advances only its own modes, then the partial products are exchanged (one all_gather per iteration)
S = orc.OracleSolve([A] * d, b, tol, nmax, orc.LANCZOS_REORTH, orc.SYM, orc.LAPLACE, tables, per_mode=True)
mine = range(first, first + count)
for k in range(2, nmax + 1):
    for s in mine:
        S._step(s, k); S.bt[s][k - 1] = S.V[s][:, k - 1] @ S.b[s]
    sc = S.schedule[k]; t = sc["t"]
    Hk = [S.H[s][:k, :k] for s in mine]; btk = [S.bt[s][:k] for s in mine]
    lam, Y = orc.solve_compressed_fast(Hk, btk, sc["alpha"], sc["omega"], sc["lambda_min"], orc.SYM, per_mode=True)
    Ly, Z, X, Lz = orc.gram_parts(Hk, Y, k)
    P0 = np.ones((t, t)); Pe = np.zeros((t, t)); Ph = np.zeros((t, t)); Peh = np.zeros((t, t)); Pg = np.zeros((t, t))
    v0 = np.ones(t); v1 = np.zeros(t); bb = 1.0
    for q, s in enumerate(mine):
        L = Ly[q]; dl = Y[q][k - 1]; g = S.H[s][k, k - 1] ** 2 * np.outer(dl, dl)
        Peh = Peh * L + Pe * X[q].T + Ph * X[q] + P0 * Lz[q]
        Pe = Pe * L + P0 * X[q]; Ph = Ph * L + P0 * X[q].T; Pg = Pg * L + P0 * g; P0 = P0 * L
        v1 = v1 * Y[q][0] + v0 * Z[q][0]; v0 = v0 * Y[q][0]; bb *= btk[q] @ btk[q]
    part = torch.from_numpy(np.concatenate([P0.ravel(), Pe.ravel(), Ph.ravel(), Peh.ravel(), Pg.ravel(), v0, v1, [bb]]))
    gathered = [torch.empty_like(part) for _ in range(2)]
    dist.all_gather(gathered, part)
    tt = t * t
    A0 = np.ones((t, t)); Ae = np.zeros((t, t)); Ah = np.zeros((t, t)); Aeh = np.zeros((t, t)); Ag = np.zeros((t, t))
    a0 = np.ones(t); a1 = np.zeros(t); abb = 1.0
    for gpart in gathered:
        p = gpart.numpy()
        B0, Be, Bh, Beh, Bg = (p[i * tt:(i + 1) * tt].reshape(t, t) for i in range(5))
        b0, b1, bbb = p[5 * tt:5 * tt + t], p[5 * tt + t:5 * tt + 2 * t], p[5 * tt + 2 * t]
        Aeh = Aeh * B0 + Ae * Bh + Ah * Be + A0 * Beh
        Ae = Ae * B0 + A0 * Be; Ah = Ah * B0 + A0 * Bh; Ag = Ag * B0 + A0 * Bg; A0 = A0 * B0
        a1 = a1 * b0 + a0 * b1; a0 = a0 * b0; abb *= bbb
    W = np.tril(2 * np.ones((t, t)), -1) + np.eye(t); Lam = np.tril(np.outer(lam, lam))
    hy2 = np.sum(W * Lam * np.tril(Aeh)); bnd = np.sum(W * Lam * np.tril(Ag)); hyb = S.b_norm * np.sum(lam * a1)
    ref = full.detail[k]
    for got, key in [(hy2, "hy2"), (bnd, "boundary"), (hyb, "hyb"), (abb, "bb")]:
        assert abs(got - ref[key]) <= 1e-12 * abs(ref[key]), (rank, k, key, got, ref[key])
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_mode_sharding_and_partial_merge_world2_gloo(tmp_path):
    """N>1 path on CPU: two ranks own disjoint mode blocks (the library's partition), exchange one partial
    product per iteration with all_gather (gloo), merge with the library's merge rule, and reproduce the
    single-process oracle.  This is the host-side contract the NCCL path implements."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm) needs no GPU: one JSON line with
    the contract's keys, the oracle port as the thing timed."""
    import json, subprocess, sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--d", "16", "--n", "500", "--nmax", "12", "--cpu-sample-modes", "8"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "krylov_iters_per_s" and d["unit"] == "iter/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_committed_ncu_launch_list_matches_its_summary():
    """profiles/: the launch list of one timed solve holds 445 launches (7 per iteration + init + the QL check) and the
    per-kernel totals in the summary are the sums of its rows."""
    import collections, csv, re
    rows = [r for r in csv.reader(open(os.path.join(ROOT, "profiles", "r01_ncu_launches_bench_n1.csv"))) if r and r[0].isdigit()]
    assert len(rows) == 445
    tot = collections.Counter()
    for r in rows:
        name = re.sub(r"\(.*$", "", r[4]).replace("void ", "").replace("tk::", "").strip()
        tot[re.sub(r"<.*$", "", name)] += float(r[-1].replace(",", "")) / (1e6 if r[-2] in ("ns", "nsecond") else 1e3)
    summary = open(os.path.join(ROOT, "profiles", "r01_ncu_launch_summary.txt")).read()
    gram = float(re.search(r"^gram_row_kernel\S*\s+\S*\s+(\d+)\s+([\d.]+)", summary, re.M).group(2))
    assert abs(gram - tot["gram_row_kernel"]) < 0.01
    assert tot["gram_row_kernel"] > tot["lanczos_ttr_bulk_kernel"] > 0


def test_argument_validation_needs_no_gpu(tk):
    """Error behaviour of the boundary: shape and enum checks (the reference's @asserts, system.jl:27-28, and its
    MethodErrors for unknown type tags) come back as TK_EINVAL / TK_EUNSUPPORTED with a message, before any CUDA
    call; a null handle is an error, never a crash."""
    import ctypes as C
    lib, capi = tk._capi.lib, tk._capi

    def create(d, ns, nmax, inst=0, cls=1, var=1, flags=1, dev=0, rank=0, world=1, uid=None):
        h = C.c_void_p()
        arr = (C.c_int64 * max(len(ns), 1))(*ns)
        rc = lib.tk_create(C.byref(h), d, arr, nmax, inst, cls, var, flags, dev, rank, world, uid)
        assert h.value is None or rc == 0
        if rc == 0:
            lib.tk_destroy(h)
        return rc, lib.tk_last_error().decode()

    EINVAL, EUNSUPPORTED = -1, -7
    assert create(0, [10], 4)[0] == EINVAL
    assert create(2, [10, 10], 0)[0] == EINVAL
    rc, msg = create(2, [10, 12], 4)
    assert rc == EUNSUPPORTED and "same order" in msg          # the reference itself only works for one n
    assert create(2, [10, 10], 11)[0] == EINVAL                  # nmax > n
    assert create(2, [0, 0], 1)[0] == EINVAL
    assert create(2, [10, 10], 4, inst=2)[0] == EINVAL
    assert create(2, [10, 10], 4, cls=6)[0] == EINVAL
    assert create(2, [10, 10], 4, var=3)[0] == EINVAL
    assert create(2, [10, 10], 4, rank=1, world=1)[0] == EINVAL
    assert create(2, [10, 10], 4, rank=0, world=2, uid=None)[0] == EINVAL
    assert lib.tk_create(None, 2, (C.c_int64 * 2)(10, 10), 4, 0, 1, 1, 1, 0, 0, 1, None) == EINVAL
    # null handle on every entry point that takes one
    z = np.zeros(8)
    i32, i64 = C.c_int32(), C.c_int64()
    assert lib.tk_solve(None, 1e-8, C.byref(i32), C.byref(i64), C.byref(i32), capi.dptr(z), capi.dptr(z), capi.dptr(z)) == EINVAL
    assert lib.tk_set_rhs(None, 0, capi.dptr(z), 8) == EINVAL
    assert lib.tk_set_rhs_all(None, capi.dptr(z), 8) == EINVAL
    assert lib.tk_set_schedule(None, 2, 1.0, 1, capi.dptr(z), capi.dptr(z)) == EINVAL
    assert lib.tk_schedule_laplace(None, 1e-8) == EINVAL
    assert lib.tk_local_modes(None, C.byref(i32), C.byref(i32)) == EINVAL
    assert lib.tk_get_solution(None, 0, capi.dptr(z), 8, capi.dptr(z), 8, 0) == EINVAL
    assert lib.tk_get_solution_all(None, capi.dptr(z), 8, capi.dptr(z), 8, 0) == EINVAL
    assert lib.tk_get_solution_device(None, capi.dptr(z), 8, None, 8, 0) == EINVAL
    assert lib.tk_get_detail(None, 2, 2, capi.dptr(z)) == EINVAL
    assert lib.tk_get_solve_info(None, None, None, None, None) == EINVAL
    assert lib.tk_launch_count(None, C.byref(i64)) == EINVAL
    assert lib.tk_begin(None) == EINVAL
    assert lib.tk_step_bases(None, 2) != 0 and lib.tk_compress(None, 2) != 0 and lib.tk_residual(None, 2, 0.0, None) != 0
    lib.tk_destroy(None)                                          # no-op
    # host-side helpers
    assert lib.tk_nonsym_coefficients(1.0, 1e-9, 3, C.byref(i32), C.byref(i32), capi.dptr(z), capi.dptr(z)) == EINVAL  # cap too small
    assert lib.tk_tables_load(b"/nonexistent/tables.bin") != 0
    capi._tables_loaded = None                                    # a failed load must not leave a stale cache marker
    capi.load_tables()
    with pytest.raises(tk.TKError):
        tk.sym_lookup(1e40, 1e-9)                                 # kappa outside the table (approximation.jl:71-76 loops forever)
    for bad in (float("inf"), float("nan"), 0.5, -3.0):
        with pytest.raises(tk.TKError) as ei:
            tk.sym_lookup(bad, 1e-9)
        assert ei.value.code == EINVAL


def test_device_code_is_the_measured_device_code(tk):
    """profiles/sass_fingerprint.txt holds one md5 per kernel (SASS instruction text) of the library the numbers under
    profiles/ were measured with.  A host-side change must leave all of them untouched; after an intended kernel
    change re-measure and refresh the record with `python tools/sass_fingerprint.py --write`."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_fingerprint as sf
    fp = sf.fingerprint(tk.LIB_PATH)
    rec = dict(reversed(l.split("  ", 1)) for l in open(sf.REC).read().splitlines() if l.strip())
    changed = sorted(k for k in set(fp) | set(rec) if fp.get(k) != rec.get(k))
    assert not changed, f"device code differs from profiles/sass_fingerprint.txt for {len(changed)} kernels, e.g. {changed[:3]}"


def test_minor_extremes_match_lapack(tk):
    """f1: the eigen-extremes of the leading minors of A_1 (eigenvalues.jl:335-350) computed inside the library
    (tk_minor_extremes: Householder + Sturm bisection / Hessenberg QR on the host threads) against LAPACK, for the
    classes the reference has a method for.  Two backward-stable algorithms agree to a few eps * ||minor||."""
    import scipy.sparse as sp
    capi, lib = tk._capi, tk._capi.lib
    eps = np.finfo(float).eps

    def ext(A, nmax, general):
        lead = np.asfortranarray(A[:nmax, :nmax].toarray() if sp.issparse(A) else np.asarray(A)[:nmax, :nmax], dtype=float)
        out = np.zeros(2 * (nmax + 1))
        assert lib.tk_minor_extremes(capi.dptr(lead), nmax, nmax, general, capi.dptr(out)) == 0, lib.tk_last_error()
        return out.reshape(-1, 2), lead

    # NonSymInstance / ConvDiff: minimum(eigvals(A_1[1:k,1:k])), an already-Hessenberg banded operator
    o, lead = ext(tk.assemble_matrix(300, tk.ConvDiff), 80, 1)
    for k in (2, 3, 17, 80):
        ev = np.linalg.eigvals(lead[:k, :k])
        assert abs(o[k, 0] - ev.real.min()) <= 200 * eps * np.abs(lead[:k, :k]).sum(axis=0).max()
        assert np.isnan(o[k, 1])
    # a dense non-symmetric matrix whose minors have real spectra (D S D^-1, S symmetric): needs the Hessenberg reduction
    rng = np.random.default_rng(0)
    X = rng.normal(size=(40, 40))
    Dg = np.exp(rng.normal(size=40))
    M = (X @ X.T + 40 * np.eye(40)) * Dg[:, None] / Dg[None, :]
    o, lead = ext(M, 30, 1)
    for k in (2, 5, 12, 30):
        ev = np.linalg.eigvals(lead[:k, :k])
        assert np.abs(ev.imag).max() == 0
        assert abs(o[k, 0] - ev.real.min()) <= 1e-12 * np.abs(ev).max()
    # SymInstance / RandSPD: both extremes of a dense symmetric minor
    S = tk.assemble_matrix(60, tk.RandSPD, rng=np.random.default_rng(3))
    o, lead = ext(S, 59, 0)
    for k in (2, 3, 30, 59):
        ev = np.linalg.eigvalsh(lead[:k, :k])
        assert abs(o[k, 0] - ev[0]) <= 100 * eps * ev[-1] and abs(o[k, 1] - ev[-1]) <= 100 * eps * ev[-1]
    # tridiagonal minors skip the reduction (the parameterised families of the experiments use the RandSPD rule)
    T = sp.diags([-np.ones(199), 1.9999 * np.ones(200), -np.ones(199)], [-1, 0, 1]).toarray()
    o, lead = ext(T, 150, 0)
    ev = np.linalg.eigvalsh(lead[:150, :150])
    assert abs(o[150, 0] - ev[0]) <= 100 * eps * 4.0 and abs(o[150, 1] - ev[-1]) <= 100 * eps * 4.0
    # a minor with complex eigenvalues: the reference's minimum(eigvals(...)) has no method there -> an error, not a number
    R = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    out = np.zeros(8)
    assert lib.tk_minor_extremes(capi.dptr(np.asfortranarray(R)), 3, 3, 1, capi.dptr(out)) == -7
    assert b"complex" in lib.tk_last_error()


def test_committed_r02_launch_list_matches_its_summary():
    """profiles/ (round 2): one whole warm solve is 383 launches -- 6 per iteration (3-term step, Gram row + monitor,
    bisection, QL fallback check, assembly + Gram blocks, combine + exchange + finalize) plus reset, init, the first
    Gram row and step 1 -- and the per-kernel totals in the summary are the sums of its rows."""
    import collections, csv, re
    rows = [r for r in csv.reader(open(os.path.join(ROOT, "profiles", "r02_ncu_launches_bench_n1.csv"))) if r and r[0].isdigit()]
    assert len(rows) == 383
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        name = re.sub(r"<.*$", "", re.sub(r"\(.*$", "", r[4]).replace("void ", "").replace("tk::", "").strip())
        tot[name] += float(r[-1].replace(",", "")) / (1e6 if r[-2] in ("ns", "nsecond") else 1e3)
        cnt[name] += 1
    assert cnt["gram_row_kernel"] == 65 and cnt["lanczos_ttr_bulk_kernel"] == 64 and cnt["combine_chunk_kernel"] == 63
    assert cnt["reset_kernel"] == 1 and "finalize_kernel" not in cnt      # the final merge runs inside the combine kernel
    summary = open(os.path.join(ROOT, "profiles", "r02_ncu_launch_summary.txt")).read()
    gram = float(re.search(r"^gram_row_kernel\s+(\d+)\s+([\d.]+)", summary, re.M).group(2))
    assert abs(gram - tot["gram_row_kernel"]) < 0.01
    assert tot["gram_row_kernel"] > tot["lanczos_ttr_bulk_kernel"] > 0


def test_committed_bench_lines_carry_the_contract_keys():
    """profiles/r02_bench_n{1,2,4,8}: one JSON line each with the contract's keys, parity green at every N."""
    import json
    for name, n in (("r02_bench_n1.json", 1), ("r02_bench_n2.json", 2), ("r02_bench_n4.json", 4), ("r02_bench_n8.json", 8)):
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        assert d["metric"] == "krylov_iters_per_s" and d["n_gpus"] == n and d["dtype"] == "f64" and d["value"] > 0
        assert d["roofline"]["bound"] == "hbm" and 0 < d["roofline"]["frac"] < 1.1 and d["gpu_launches"] > 0
        assert d["parity"]["checked"] and d["parity"]["ok"]
        assert "workload" in d["config"] and d["clocks"]["sm_mhz"] > 0
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_n1.json")))
    assert d["cpu_baseline"]["kind"] == "port" and "nothing scaled" in d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["time_to_tol"]["with_solution_ms"] > d["time_to_tol"]["device_ms"]
    # the other Lanczos variant at the same sizes rides on the default line (no Gram row over all modes: faster)
    assert d["variants"]["TensorLanczos"]["value"] > d["value"]
