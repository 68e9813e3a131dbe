#!/bin/bash
# kernel tuning sweep (developer tool): single stream so the per-kernel timings are uncontended
export TK_SINGLE_STREAM=1
for cfg in "TK_TTR_DIA=0" "TK_TTR_CPM=2" "TK_TTR_CPM=4" "TK_TTR_CPM=4 TK_TTR_THREADS=512" "TK_TTR_CPM=8" "TK_TTR_CPM=8 TK_TTR_THREADS=256" "TK_TTR_CPM=1"; do
  echo "== $cfg"; env $cfg python tools/profile_phases.py 1024 10000 64 reorth | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('ttr','gram','solve')})"
done
