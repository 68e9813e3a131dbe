#!/bin/bash
# developer tool: 3-term step, bulk-copy kernel vs the load/store kernel, cluster width (C5, one GPU)
D=${1:-1024}
for cfg in "TK_TTR_BULK=0" "TK_TTR_BULK=1" "TK_TTR_CPM=2" "TK_TTR_CPM=8" "TK_TTR_NOCONST=1"; do
  echo "== d=$D $cfg"
  env $cfg timeout 300 python tools/profile_phases.py $D 10000 64 reorth | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('ttr','gram','solve','relres_last')})"
done
