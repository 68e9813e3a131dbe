#!/bin/bash
# GPU call 6 (1 GPU): programmatic dependent launch on the Krylov-step stream, A/B at 1024 / 128 / 32 modes; tests under PDL.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
rm -f $O/*.ncu-rep $O/pdl.log
for d in 1024 128 32; do
  for cfg in "TK_PDL=0" "TK_PDL=1"; do
    echo "== d=$d $cfg" >> $O/pdl.log
    env $cfg timeout 300 python bench.py --d $d --steps 10 --warmup 3 --no-extras --no-cpu-baseline >> $O/pdl.log 2>> $O/pdl.err
  done
done
( time TK_PDL=1 timeout 1800 python -m pytest tests -m gpu -q -x ) > $O/pytest_gpu_pdl.log 2>&1
tail -5 $O/pytest_gpu_pdl.log
( time timeout 1800 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log
