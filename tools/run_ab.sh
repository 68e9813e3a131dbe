#!/bin/bash
# A/B of one environment knob on the default bench: tools/run_ab.sh KNOB v1 v2 [extra bench args]
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
knob=$1; v1=$2; v2=$3; shift 3
rm -f $O/ab_*.jsonl
for v in $v1 $v2 $v1 $v2; do
  env $knob=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras "$@" >> $O/ab_$v.jsonl 2>> $O/ab.err
done
python - $v1 $v2 <<'PY'
import json, sys
for v in sys.argv[1:]:
    for l in open(f"gpurun_out/ab_{v}.jsonl"):
        r = json.loads(l); rf = r["roofline"]
        print(v, round(r["value"], 1), r["ms_per_step"], round(rf["frac"], 3), round(rf["avg_launch_ms"], 4), rf["ttr_kernel"], (r.get("parity") or {}).get("ok"))
PY
