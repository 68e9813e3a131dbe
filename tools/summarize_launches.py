#!/usr/bin/env python3
"""ncu launch list (`--metrics gpu__time_duration.sum`) of `bench.py --steps 1 --warmup 3 --no-extras` ->
profiles/<tag>_ncu_launch_summary.txt and profiles/<tag>_ncu_launches_bench_n1.csv (ONE whole solve: the launches
between two reset_kernel launches, taken from a warm solve).

    python tools/summarize_launches.py gpurun_out/r02_launches_bench_n1.csv r02
"""
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, tag = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "r02")
raw = list(csv.reader(open(src)))
hdr = next(r for r in raw if r and r[0] == "ID")
rows = [r for r in raw if r and r[0].isdigit()]


def short(name):
    name = re.sub(r"\(.*$", "", name).replace("void ", "").replace("tk::", "").strip()
    return re.sub(r"<.*$", "", name)


starts = [i for i, r in enumerate(rows) if short(r[4]) == "reset_kernel"]
assert len(starts) >= 3, "need at least two whole solves in the list"
a, b = starts[1], starts[2]
solve = rows[a:b]
tot = collections.OrderedDict()
for r in solve:
    name, unit, val = short(r[4]), r[-2], float(r[-1].replace(",", ""))
    ms = val / 1e6 if unit in ("ns", "nsecond") else val / 1e3 if unit in ("us", "usecond") else val
    n, t = tot.get(name, (0, 0.0))
    tot[name] = (n + 1, t + ms)
total = sum(t for _, t in tot.values())
lines = [f"# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline",
         f"# window = one whole warm solve (C5: d=1024, n=10^4, nmax=64): the {len(solve)} launches between two reset_kernel launches;",
         "# times are cold-cache and serialised by the profiler: compare SHARES, not absolutes",
         f"{'kernel':<44} {'launches':>8} {'total ms':>10} {'share':>7}"]
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{name:<44} {n:>8} {t:>10.3f} {100 * t / total:>6.1f}%")
lines.append(f"{'total':<44} {len(solve):>8} {total:>10.3f}")
kry = sum(t for k, (n, t) in tot.items() if k.startswith(("gram_row", "lanczos_ttr", "init_basis", "arnoldi", "reset")))
top = sum(t for k, (n, t) in tot.items() if k.startswith("gram_row"))
lines.append(f"# Krylov-step stream (reset + init + 3-term step + Gram row incl. monitor): {kry:.3f} ms; Gram row share of that stream: {100 * top / kry:.1f}%")
lines.append("# The eigensolver / assembly / combine kernels run on side streams concurrently with that stream in the live run,")
lines.append("# so the step time IS that stream; bench.py reports the live shares (roofline.share_of_step, ttr_kernel.share_of_step).")
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_launch_summary.txt"), "w").write("\n".join(lines) + "\n")
with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_launches_bench_n1.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(hdr)
    w.writerows(solve)
print("\n".join(lines))
