#!/bin/bash
# GPU call 9 (1 GPU): the new vector of the Gram row from L1 instead of staged shared memory; 3-term cluster width at
# small mode counts; Gram grid at 512 modes.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
rm -f $O/ab2.log
for d in 128 1024 32; do
  for cfg in "TK_GRAM_MODE=0" "TK_GRAM_MODE=0 TK_GRAM_WSMEM=0" "TK_GRAM_MODE=0 TK_GRAM_WSMEM=0 TK_GRAM_CPC=64"; do
    echo "== d=$d $cfg" >> $O/ab2.log
    env $cfg timeout 300 python bench.py --d $d --steps 10 --warmup 3 --no-extras --no-cpu-baseline >> $O/ab2.log 2>> $O/ab2.err
  done
done
for d in 128 32; do
  for cfg in "TK_TTR_CPM=8" "TK_TTR_CPM=2" "TK_TTR_CPM=4"; do
    echo "== d=$d $cfg" >> $O/ab2.log
    env $cfg timeout 300 python bench.py --d $d --steps 10 --warmup 3 --no-extras --no-cpu-baseline >> $O/ab2.log 2>> $O/ab2.err
  done
done
for cfg in "TK_GRAM_MODE=0" "TK_GRAM_MODE=1"; do
  echo "== d=512 $cfg" >> $O/ab2.log
  env $cfg timeout 300 python bench.py --d 512 --steps 10 --warmup 3 --no-extras --no-cpu-baseline >> $O/ab2.log 2>> $O/ab2.err
done
echo "== d=128 nmax=256 TK_GRAM_MODE=0 vs 1" >> $O/ab2.log
TK_GRAM_MODE=0 timeout 300 python bench.py --d 128 --nmax 256 --steps 3 --warmup 3 --no-extras --no-cpu-baseline >> $O/ab2.log 2>> $O/ab2.err
TK_GRAM_MODE=1 timeout 300 python bench.py --d 128 --nmax 256 --steps 3 --warmup 3 --no-extras --no-cpu-baseline >> $O/ab2.log 2>> $O/ab2.err
