#!/usr/bin/env python3
"""Condense `ncu -i X.ncu-rep --page raw --csv` exports (gpurun_out/*.raw.csv) into one tracked text table:

    python tools/summarize_ncu_raw.py gpurun_out/r02_c5_side.raw.csv [more.csv ...] > profiles/r02_ncu_secondary_kernels.txt
"""
import csv
import re
import sys

UNIT = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6,   # -> us
        "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}                                                        # -> MB

KEEP = [("gpu__time_duration.sum", "dur_us", None),
        ("dram__bytes_read.sum", "dram_rd_MB", None), ("dram__bytes_write.sum", "dram_wr_MB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%", 1),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%", 1),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64pipe_%", 1),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64act_%", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%", 1),
        ("launch__registers_per_thread", "regs", 1), ("launch__grid_size", "grid", 1), ("launch__block_size", "block", 1),
        ("launch__cluster_size", "cluster", 1),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long", 1),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar", 1),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short", 1),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st_membar", 1),
        ("lts__t_sector_hit_rate.pct", "l2hit_%", 1)]


def short(name):
    name = re.sub(r"\(.*$", "", name).replace("void ", "").replace("tk::", "").strip()
    return name[:46]


def main():
    print("# from `ncu --set full --clock-control none` captures (tools/run_evidence_1gpu.sh); one row per profiled launch")
    hdrline = f"{'kernel':<46} " + " ".join(f"{lab:>10}" for _, lab, _ in KEEP)
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        hdr = rows[hi]
        units = rows[hi + 1]
        col = {name: hdr.index(name) for name, _, _ in KEEP if name in hdr}
        kn = hdr.index("Kernel Name")
        print(f"\n## {path.split('/')[-1]}")
        print(hdrline)
        for r in rows[hi + 2:]:
            if not r or not r[0].isdigit():
                continue
            vals = []
            for name, lab, sc in KEEP:
                if name in col:
                    try:
                        v = float(r[col[name]].replace(",", "")) * (UNIT.get(units[col[name]], 1.0) if sc is None else sc)
                        vals.append(f"{v:>10.3g}" if abs(v) < 1e5 else f"{v:>10.0f}")
                    except ValueError:
                        vals.append(f"{'-':>10}")
                else:
                    vals.append(f"{'n/a':>10}")
            print(f"{short(r[kn]):<46} " + " ".join(vals))


if __name__ == "__main__":
    main()
