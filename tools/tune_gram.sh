#!/bin/bash
# Gram-row kernel knobs at 1024 and 128 modes per GPU (chunks grid, new vector through L1).
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
: > $O/tune_gram.log
for d in 1024 128; do
  for cfg in "TK_GRAM_U=4" "TK_GRAM_U=8" "TK_GRAM_U=2" "TK_GRAM_THREADS=512" "TK_GRAM_CPC=16" "TK_GRAM_CPC=64" "TK_GRAM_CPC=48 TK_GRAM_U=8" "TK_GRAM_WPC=4" ; do
    echo "== d=$d $cfg" >> $O/tune_gram.log
    env $cfg timeout 300 python bench.py --d $d --steps 8 --warmup 3 --no-extras --no-cpu-baseline >> $O/tune_gram.log 2>> $O/tune_gram.err
  done
done
