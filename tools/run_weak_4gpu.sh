#!/bin/bash
# 4-GPU box: weak scaling (128 modes per GPU) at N = 4 and 2.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
P=29927
for n in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --weak --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_weak_n$n.json 2> $O/bench_weak_n$n.err
  P=$((P+1))
  head -c 200 $O/bench_weak_n$n.json; echo
done
