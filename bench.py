#!/usr/bin/env python3
"""bench.py -- Krylov iterations/s of the tensorized solve on B200 (BASELINE.json metric).

Workload (config 5 of BASELINE.json): d = 1024 modes, n_s = 10^4, 1D Laplacian per mode,
rank-1 right-hand side (one U(0,1) vector for all modes, seed 12345, normalised),
TensorLanczosReorth, exp-sum schedule from the coefficient tables at tol 1e-8, reference
semantics (exp(gamma H_1) for all modes), fixed-iteration mode (nmax-1 Krylov iterations per
solve; BASELINE.md: no reference run reaches its tolerance, throughput is quoted per iteration).

A "step" is one whole solve.  value = Krylov iterations (all d modes advanced) per second, timed
with CUDA events on the library's stream, max over ranks.  With --gpus N the d modes are
block-partitioned over N ranks (strong scaling), one NCCL all-gather + one 3*(nmax+1)-double
broadcast per iteration.

`--impl reference` times the CPU oracle (the reference is Julia, which this image does not have)
on the host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

METRIC = "krylov_iters_per_s"
UNIT = "iter/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--d", type=int, default=1024)
    ap.add_argument("--n", type=int, default=10000)
    ap.add_argument("--nmax", type=int, default=64)
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--variant", default="reorth", choices=["reorth", "lanczos"])
    ap.add_argument("--per-mode", action="store_true", help="each mode exponentiates its own H_s (not the reference's H_1)")
    ap.add_argument("--cpu-sample-modes", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(a):
    tag = "C5" if (a.d, a.n, a.nmax) == (1024, 10000, 64) else "custom"
    return (f"{tag} d={a.d} n={a.n} Laplace TensorLanczos{'Reorth' if a.variant == 'reorth' else ''} nmax={a.nmax} "
            f"tol={a.tol:g} fixed-iterations {'per-mode H_s' if a.per_mode else 'reference H_1'}")


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md), sampled in-process through NVML
    (a 20 ms thread; spawning `nvidia-smi -lms` perturbs short multi-GPU steps), nvidia-smi as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread, self.stop_flag = index, [], None, None, False
        self.sm, self.smax, self.reasons = [], [], set()

    def _nvml_loop(self, nv, handle):
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        try:
            smax = nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM)
        except Exception:
            smax = None
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)))
                if smax:
                    self.smax.append(float(smax))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                for name, bit in bits.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.02)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            handle = nv.nvmlDeviceGetHandleByIndex(phys)
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
        except OSError:
            self.proc = None
            return
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.thread:
            self.stop_flag = True
            self.thread.join(timeout=1)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None,
                    "sm_max_mhz": max(self.smax) if self.smax else None, "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "source": "nvml"}
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def measured_traffic_ratio():
    """DRAM bytes / algorithmic bytes of the Gram-row kernel from the committed `ncu --set full` capture."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_gram_traffic.json")))
        return float(t["traffic_over_algorithmic"]), "profiles/r01_gram_traffic.json (ncu dram__bytes_read+write, launches with 31-33 columns)"
    except Exception:
        return None, None


def measured_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores, bounded sample of the same workload
# ---------------------------------------------------------------------------------------------
def cpu_arm(a, sample_modes):
    """Times the oracle (flavour B of BASELINE.md: one eigendecomposition per iteration and the O(d t^2) combine,
    i.e. the GPU path's algorithm; the reference's own O(d^3 t^2) loops cannot run at d = 1024) on
    `sample_modes` of the d modes for all nmax-1 iterations.  Per-iteration cost is linear in the number of
    modes, so iterations/s at d modes = (nmax-1) / (t_sample * d / sample_modes)."""
    orc = entry.load_oracle()
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank)
    threads = os.cpu_count()
    tk_tables = os.path.join(entry.PKG_DIR, "data", "expsum_tables.bin")
    tables = orc.ExpSumTables.from_packed(tk_tables)
    ds = min(sample_modes, a.d)
    A = orc.assemble_matrix(a.n, orc.LAPLACE)
    b = orc.normalize_rhs(orc.random_rhs(ds, a.n, 12345))
    sched = orc.build_schedule(A, a.d, a.nmax, a.tol, orc.SYM, orc.LAPLACE, tables)
    variant = orc.LANCZOS_REORTH if a.variant == "reorth" else orc.LANCZOS
    t0 = time.perf_counter()
    # one worker thread per core over the (independent) modes, BLAS itself single-threaded inside each worker
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass
    S = orc.OracleSolve([A] * ds, b, a.tol, a.nmax, variant, orc.SYM, orc.LAPLACE, tables, per_mode=a.per_mode,
                        residual="nilpotent", fast_solve=True, schedule=sched, ignore_breakdown=True,
                        mode_threads=threads)
    S.run()
    dt = time.perf_counter() - t0
    its = (a.nmax - 1) / (dt * a.d / ds)
    return {"value": its, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{ds} of {a.d} modes x {a.nmax - 1} iterations in {dt:.1f} s, scaled linearly in d; numpy/scipy, "
                      f"{threads} worker threads over the modes; oracle flavour B (GPU-matched algorithm)"}, dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    vals, dts = [], []
    for _ in range(max(a.warmup, 0) and 1):
        cpu_arm(a, max(4, a.cpu_sample_modes // 8))
    for _ in range(max(a.steps, 1)):
        cb, dt = cpu_arm(a, a.cpu_sample_modes)
        vals.append(cb["value"]); dts.append(dt)
        if sum(dts) > 150:
            break
    v = float(np.mean(vals))
    cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": len(vals),
            "warmup": a.warmup, "ms_per_step": 1e3 * (a.nmax - 1) / v, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "note": "Julia is not installed; the CPU arm is the oracle port"},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_b200(a):
    # Only the one JSON line may reach stdout (NCCL prints its version banner there): park fd 1 on stderr until
    # the line is ready.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    tk = entry.load_package()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            import ctypes as C
            raw = C.create_string_buffer(128)
            tk._capi.check(tk._capi.lib.tk_comm_unique_id(raw))
            buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())

    d, n, nmax = a.d, a.n, a.nmax
    variant = tk.TensorLanczosReorth if a.variant == "reorth" else tk.TensorLanczos
    base = tk.TK_FLAG_FIXED_ITERATIONS | (0 if a.per_mode else tk.TK_FLAG_REFERENCE_H1)
    A1 = tk.assemble_matrix(n, tk.Laplace)
    # pinned host buffers: the step's inputs
    b_host = torch.from_numpy(np.random.default_rng(12345).random(n)).pin_memory()
    b_np = b_host.numpy()
    b_np *= 1.0 / np.linalg.norm(b_np)        # TensorizedSystem normalises b (system.jl:33-37)

    def make(flags):
        s = tk.Solver(d, n, nmax, tk.SymInstance, tk.Laplace, variant, flags=flags, device=local, rank=rank,
                      world=world, unique_id=uid)
        return s

    def feed(s):
        s.set_operators([A1] * d)
        s.set_rhs([b_np] * d)
        s.set_schedule(A1, a.tol)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm: inputs already in HBM, handle reused; kernel events on -------
    slv = make(base | tk.TK_FLAG_TIME_KERNELS)
    feed(slv)
    for _ in range(a.warmup):
        slv.solve(a.tol)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    dev_ms, gram_ms, gram_bytes, gram_n, ttr_ms, ttr_bytes, launches = 0.0, 0.0, 0.0, 0, 0.0, 0.0, 0
    slv.timing_mark()               # CUDA event on the library's stream: start of the K-step timed region
    t0 = time.perf_counter()
    for _ in range(a.steps):
        res = slv.solve(a.tol)
        ms, cnt, by = slv.timing(1)
        gram_ms += ms; gram_n += cnt; gram_bytes += by
        ms, cnt, by = slv.timing(0)
        ttr_ms += ms; ttr_bytes += by
        launches += slv.launch_count()
    dev_ms = slv.timing(7)[0]       # region start -> end of the last solve, on the device
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = maxr(dev_ms)
    gram_ms = maxr(gram_ms)
    relres_last = float(res["relres"][nmax - 1])
    status = res["status"]
    fb = slv.orth_state(slv.first)[1] if slv.count else 0
    slv.close()

    # ---- time-to-tolerance (second half of BASELINE.json's metric): parity mode, tol 1e-5 -- the only tolerance
    # this configuration reaches: device time from first enqueue to the converged status
    ttt = None
    if a.d >= 256 and a.n >= 10000:
        s2 = make(0 if a.per_mode else tk.TK_FLAG_REFERENCE_H1)
        s2.set_operators([A1] * d); s2.set_rhs([b_np] * d); s2.set_schedule(A1, 1e-5)
        best = None
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            r2 = s2.solve(1e-5)
            wall = 1e3 * (time.perf_counter() - t0)
            dev = maxr(s2.timing(6)[0])
            if best is None or dev < best[0]:
                best = (dev, wall)
        ttt = {"tol": 1e-5, "status": r2["status"], "iterations": r2["term_k"], "device_ms": best[0], "call_ms": best[1],
               "relres": float(r2["relres"][r2["term_k"] - 1]) if r2["term_k"] >= 1 else None}
        s2.close()

    # ---- end-to-end arm: the public call with host buffers, H2D and D2H inside the timed region -----
    A = tk.KroneckerMatrix(tk.SymInstance, [A1] * d, tk.Laplace)
    def e2e_once():
        out = []
        cd = tk.ConvergenceData(nmax)
        s = make(base)
        try:
            feed(s)
            r = s.solve(a.tol)
            cd.relative_residual_norm[:] = r["relres"]
        finally:
            s.close()
        return cd
    for _ in range(min(a.warmup, 2)):
        e2e_once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        cd = e2e_once()
    barrier()
    e2e_ms = maxr(1e3 * (time.perf_counter() - t0))
    assert abs(cd.relative_residual_norm[nmax - 1] - relres_last) <= 1e-9 * abs(relres_last) + 1e-300
    dl = slv.count
    # bytes that actually cross PCIe per call: the reference aliases ONE rhs vector and ONE matrix over all d modes
    # (system.jl:5-11, tensor_struct.jl:208-210) and the library honours that (tk_set_rhs_all, tk_share_operator):
    # b once, the CSC arrays of A_1 once (colptr, rowval, nzval), the exp-sum schedule (<= 64 terms per iteration)
    h2d = n * 8 + (n + 1) * 8 + 2 * int(A1.nnz) * 8 + sum(16 * 64 for _ in range(nmax))
    d2h = 3 * nmax * 8 + 64

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    def emit(line):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)

    iters = a.steps * (nmax - 1)
    value = iters / (dev_ms / 1e3)
    peak, peak_kind = measured_peak()
    achieved = gram_bytes / (gram_ms / 1e3) / 1e9 if gram_ms > 0 else 0.0
    tr_ratio, tr_src = measured_traffic_ratio()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "iterations_per_step": nmax - 1, "modes_per_gpu": dl,
                   "l2": "inputs larger than L2: the Krylov bases are %.1f GB per GPU, re-streamed every iteration"
                         % (dl * n * (nmax + 1) * 8 / 1e9),
                   "status": status, "relres_last": relres_last, "mgs_fallbacks_mode0": fb,
                   "wall_ms_per_step": wall_ms / a.steps},
        "clocks": clocks,
        "e2e": {"value": iters / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / a.steps,
                "note": "handle creation (cudaMalloc of the bases), operator/rhs/schedule upload, solve, histories back"},
        "gpu_launches": launches,
        "time_to_tol": ttt,
        "roofline": {"kernel": "gram_row_kernel (orthogonality monitor of the batched Lanczos step)", "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback",
                     "frac_of_nominal_8TBs": achieved / 8000.0, "launches": gram_n,
                     "algorithmic_bytes_per_launch_avg": gram_bytes / max(gram_n, 1),
                     "avg_launch_ms": gram_ms / max(gram_n, 1), "share_of_step": gram_ms / dev_ms,
                     "traffic": (tr_ratio * gram_bytes / max(gram_n, 1)) if (tr_ratio and world == 1) else None,
                     "traffic_source": tr_src,
                     "ttr_kernel": {"achieved": (ttr_bytes / (ttr_ms / 1e3) / 1e9) if ttr_ms > 0 else None,
                                    "share_of_step": ttr_ms / dev_ms}},
    }
    if world == 1 and not a.no_cpu_baseline:
        cb, _ = cpu_arm(a, a.cpu_sample_modes)
        line["cpu_baseline"] = cb
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    return run_b200(a)


if __name__ == "__main__":
    sys.exit(main())
