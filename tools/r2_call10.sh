#!/bin/bash
# GPU call 10 (1 GPU): Gram-row variants across mode counts: (chunks grid, staged) / (chunks grid, L1) / balanced.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
rm -f $O/ab3.log
for d in 32 64 128 256 512 1024; do
  for cfg in "TK_GRAM_MODE=0 TK_GRAM_WSMEM=1" "TK_GRAM_MODE=0 TK_GRAM_WSMEM=0" "TK_GRAM_MODE=1 TK_GRAM_WSMEM=1"; do
    echo "== d=$d $cfg" >> $O/ab3.log
    env $cfg timeout 300 python bench.py --d $d --steps 10 --warmup 3 --no-extras --no-cpu-baseline >> $O/ab3.log 2>> $O/ab3.err
  done
done
