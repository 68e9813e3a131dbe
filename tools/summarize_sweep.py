#!/usr/bin/env python3
"""gpurun_out/r02_c5_sweep.jsonl (tools/run_sweeps.sh) -> profiles/r02_c5_sweep.json: one row per sweep point."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r02_c5_sweep.jsonl")
rows = []
for line in open(src):
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d["roofline"]
    rows.append({"workload": d["config"]["workload"], "n_gpus": d["n_gpus"], "iters_per_s": round(d["value"], 1),
                 "ms_per_solve": round(d["ms_per_step"], 3), "iterations_per_solve": d["config"]["iterations_per_step"],
                 "top_kernel_frac_of_measured_hbm_peak": round(r["frac"], 3), "top_kernel_share_of_step": round(r["share_of_step"], 3),
                 "three_term_kernel_GBs": round(r["ttr_kernel"]["achieved"] or 0.0, 1),
                 "all_krylov_bytes_over_step_GBs": round(r["all_krylov_bytes_over_step"], 1),
                 "relres_last": d["config"]["relres_last"], "sm_mhz": (d.get("clocks") or {}).get("sm_mhz"),
                 "throttle": (d.get("clocks") or {}).get("reasons")})
out = os.path.join(ROOT, "profiles", "r02_c5_sweep.json")
json.dump({"source": "tools/run_sweeps.sh on one B200; bench.py --no-extras (device-resident arm, CUDA-graph replay)",
           "rows": rows}, open(out, "w"), indent=1)
print(f"{len(rows)} rows -> {out}")
