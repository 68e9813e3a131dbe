#!/bin/bash
# kernel tuning sweep (developer tool): 3-term step cluster width / threads at the per-GPU size of the 8-GPU run
for cfg in "TK_TTR_CPM=4" "TK_TTR_CPM=2" "TK_TTR_CPM=1" "TK_TTR_CPM=8" "TK_TTR_CPM=4 TK_TTR_THREADS=256" "TK_TTR_CPM=8 TK_TTR_THREADS=256" "TK_TTR_CPM=2 TK_TTR_THREADS=1024"; do
  echo "== d=128 $cfg"; env $cfg python tools/profile_phases.py 128 10000 64 reorth | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('ttr','gram','solve')})"
done
