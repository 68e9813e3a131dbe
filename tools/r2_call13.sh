#!/bin/bash
# GPU call 13 (1 GPU): the fused per-mode Lanczos step: tests, A/B at 128 / 100 / 148 modes, nmax 64 and 256; fallback cost.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
rm -f $O/fused.log
( time timeout 1800 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1
tail -6 $O/pytest_gpu.log
for d in 128 148 96; do
  for cfg in "TK_FUSED_STEP=0" "TK_FUSED_STEP=1"; do
    echo "== d=$d $cfg" >> $O/fused.log
    env $cfg timeout 300 python bench.py --d $d --steps 10 --warmup 3 --no-extras --no-cpu-baseline >> $O/fused.log 2>> $O/fused.err
  done
done
for cfg in "TK_FUSED_STEP=0" "TK_FUSED_STEP=1"; do
  echo "== d=128 nmax=256 $cfg" >> $O/fused.log
  env $cfg timeout 300 python bench.py --d 128 --nmax 256 --steps 3 --warmup 3 --no-extras --no-cpu-baseline >> $O/fused.log 2>> $O/fused.err
done
echo "== d=64 TK_FUSED_STEP=1 (TK_FUSED_MIN ignored when forced)" >> $O/fused.log
TK_FUSED_STEP=1 timeout 300 python bench.py --d 64 --steps 10 --warmup 3 --no-extras --no-cpu-baseline >> $O/fused.log 2>> $O/fused.err
echo "== C2 fused" >> $O/fused.log
TK_FUSED_STEP=1 timeout 300 python bench.py --config C2 --steps 5 --warmup 3 --no-extras --no-cpu-baseline >> $O/fused.log 2>> $O/fused.err
timeout 900 python tools/fallback_report.py > $O/fallback.log 2>&1
tail -3 $O/fallback.log
