#!/bin/bash
# several settings of the environment in one call: tools/run_ab_multi.sh "A=1 B=2" "A=3" ... (first entry "-" = defaults)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
rm -f $O/abm.jsonl $O/abm.err
for rep in 1 2; do
for setting in "$@"; do
  [ "$setting" = "-" ] && setting="TK_NOOP=0"
  env $setting timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>> $O/abm.err | python -c "
import json, sys
r = json.loads(sys.stdin.read()); rf = r['roofline']
print('$setting', round(r['value'], 1), r['ms_per_step'], round(rf['frac'], 3), round(rf['avg_launch_ms'], 4), round(rf['ttr_kernel']['achieved'], 1), round(rf['ttr_kernel']['share_of_step'], 4), (r.get('parity') or {}).get('ok'))
" | tee -a $O/abm.jsonl
done
done
