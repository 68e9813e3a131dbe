#!/usr/bin/env python3
"""Turn the files tools/run_ncu.sh leaves in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_ncu.py [round-tag]        (default r01)

Reads gpurun_out/launches.csv (the --metrics gpu__time_duration.sum pass) and gpurun_out/prof_{gram,ttr}.ncu-rep
(the two --set full captures; needs `ncu` on PATH to export them)."""
import collections, csv, io, json, os, re, subprocess, sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct", "dram__bytes_read.sum.per_second",
        "launch__shared_mem_per_block_dynamic"]


def short(name):
    name = re.sub(r"\(.*$", "", name).replace("void ", "").replace("tk::", "").strip()
    return re.sub(r"\(int\)|\(bool\)", "", name)


def launch_summary():
    rows = [r for r in csv.reader(open(os.path.join(OUT, "launches.csv"))) if r and r[0].isdigit()]
    # columns: ID, Process ID, Process Name, Host Name, Kernel Name, ..., Metric Name, Metric Unit, Metric Value
    tot = collections.OrderedDict()
    for r in rows:
        name, unit, val = short(r[4]), r[-2], float(r[-1].replace(",", ""))
        ms = val / 1e6 if unit in ("ns", "nsecond") else val / 1e3 if unit in ("us", "usecond") else val
        n, t = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, t + ms)
    total = sum(t for _, t in tot.values())
    lines = ["# ncu --metrics gpu__time_duration.sum --clock-control none -s 445 -c 445 python bench.py --steps 1 --warmup 1 --no-cpu-baseline",
             "# window = one timed solve (C5: d=1024, n=10^4, nmax=64, 445 launches); times are cold-cache and serialised",
             f"{'kernel':<60} {'launches':>8} {'total ms':>10} {'share':>7}"]
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{name:<60} {n:>8} {t:>10.3f} {100 * t / total:>6.1f}%")
    lines.append(f"{'total':<60} {sum(n for n, _ in tot.values()):>8} {total:>10.3f}")
    krylov = sum(t for k, (n, t) in tot.items() if k.startswith(("gram_row", "lanczos_ttr", "init_basis", "arnoldi")))
    gram = sum(t for k, (n, t) in tot.items() if k.startswith("gram_row"))
    lines.append(f"# Krylov-step stream (gram_row [incl. monitor] + lanczos_ttr + init): {krylov:.3f} ms; "
                 f"gram_row share of that stream: {100 * gram / krylov:.1f}%")
    lines.append("# In the live run the eigensolver / assembly / residual kernels run on other streams concurrently with this")
    lines.append("# stream, so the step time is the Krylov-step stream; bench.py reports the live shares (roofline.share_of_step).")
    open(os.path.join(PROF, f"{tag}_ncu_launch_summary.txt"), "w").write("\n".join(lines) + "\n")
    with open(os.path.join(PROF, f"{tag}_ncu_launches_bench_n1.csv"), "w") as f:
        f.write(open(os.path.join(OUT, "launches.csv")).read())
    print("\n".join(lines))


def full(rep, outname):
    raw = subprocess.run(["ncu", "-i", os.path.join(OUT, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    text, recs = [], []
    for r in rows[2:]:
        text.append("---")
        text.append(f"  Kernel Name: {r[hdr.index('Kernel Name')]} ")
        rec = {}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                text.append(f"  {k}: {r[i]} {units[i]}")
                rec[k] = (float(r[i].replace(",", "")), units[i])
        recs.append(rec)
    open(os.path.join(PROF, outname), "w").write("\n".join(text) + "\n")
    return recs


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


if __name__ == "__main__":
    launch_summary()
    recs = full("prof_gram.ncu-rep", f"{tag}_ncu_gram_row_kernel_full.txt")
    # launches 95..97 of gram_row in the bench = the timed solve's launches with 31..33 columns (65 per solve:
    # the first solve's 65 + 30 of the second are skipped)
    ncols = [31, 32, 33][:len(recs)]
    dram = [to_bytes(*r["dram__bytes_read.sum"]) + to_bytes(*r["dram__bytes_write.sum"]) for r in recs]
    alg = [8 * 10000 * c * 1024 for c in ncols]
    json.dump({"kernel": "gram_row_kernel<4,256>",
               "source": "ncu --set full --clock-control none -k regex:gram_row -s 95 -c 3 python bench.py --steps 1 --warmup 1 "
                         f"--no-cpu-baseline (profiles/{tag}_ncu_gram_row_kernel_full.txt)",
               "ncols": ncols, "dram_bytes": dram, "algorithmic_bytes": alg,
               "traffic_over_algorithmic": sum(dram) / sum(alg)},
              open(os.path.join(PROF, f"{tag}_gram_traffic.json"), "w"), indent=1)
    print("gram traffic / algorithmic:", sum(dram) / sum(alg))
    full("prof_ttr.ncu-rep", f"{tag}_ncu_lanczos_ttr_kernel_full.txt")
