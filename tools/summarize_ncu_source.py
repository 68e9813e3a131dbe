#!/usr/bin/env python3
"""`ncu -i X.ncu-rep --page source --csv` -> stall-reason totals and the hottest instructions of the first kernel.

    python tools/summarize_ncu_source.py gpurun_out/r02_c5_ttr.source.csv > profiles/r02_ncu_ttr_bulk_stalls.txt
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, tables = None, []
for r in rows:
    if r and r[0] == "Kernel Name":
        tables.append({"name": r[1], "rows": []}); hdr = None; continue
    if r and r[0] == "Address":
        hdr = r; tables[-1]["hdr"] = r; continue
    if hdr and r and r[0].startswith("0x"):
        tables[-1]["rows"].append(r)
t = tables[0]
h = t["hdr"]
isamp, iexec = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[isamp]) for r in t["rows"])
print(f"# {t['name']}")
print(f"# {len(t['rows'])} SASS instructions, {tot} warp-stall samples (ncu --set full --import-source on, source page)")
print("\nstall reason             samples  share")
agg = {s: sum(int(r[h.index(s)]) for r in t["rows"]) for s in stalls}
for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print(f"{s:24s} {v:7d} {100 * v / tot:5.1f}%")
print("\nhottest instructions: samples, share, times executed (warp level), instruction, dominant stall")
for r in sorted(t["rows"], key=lambda r: -int(r[isamp]))[:30]:
    top = max((int(r[h.index(s)]), s) for s in stalls)
    print(f"{int(r[isamp]):6d} {100 * int(r[isamp]) / tot:5.1f}% {r[iexec]:>8s}  {' '.join(r[1].split())[:64]:64s} {top[1]}")
