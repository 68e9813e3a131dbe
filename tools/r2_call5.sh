#!/bin/bash
# GPU call 5 (1 GPU): tests with the row-sliced Gram kernel and the batched SpMV, A/B at 1024 / 256 / 128 / 32 modes, C4.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
rm -f $O/*.ncu-rep $O/ab.log $O/configs.jsonl $O/d128.log
( time timeout 1800 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1
tail -8 $O/pytest_gpu.log
for d in 1024 256 128 32; do
  for cfg in "TK_GRAM_MODE=-1" "TK_GRAM_MODE=0" "TK_GRAM_MODE=1" "TK_GRAM_MODE=2"; do
    echo "== d=$d $cfg" >> $O/d128.log
    env $cfg timeout 300 python bench.py --d $d --steps 10 --warmup 3 --no-extras --no-cpu-baseline >> $O/d128.log 2>> $O/d128.err
  done
done
for cfg in "TK_BGS_THREADS=512" "TK_BGS_THREADS=256" "TK_BGS_THREADS=1024" "TK_MGS_BLOCK=0"; do
  echo "== C4 $cfg" >> $O/ab.log
  env $cfg timeout 300 python bench.py --config C4 --steps 5 --warmup 3 --no-extras --no-cpu-baseline >> $O/ab.log 2>> $O/ab.err
done
for c in C1 C2; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-extras --no-cpu-baseline >> $O/configs.jsonl 2>> $O/configs.err
done
du -sh $O
