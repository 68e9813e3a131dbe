#!/bin/bash
# 8-GPU box, short form: config 5 at N = 8, 4, 2 (default bench lines).
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
P=29917
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_n$n.json 2> $O/bench_n$n.err
  P=$((P+1))
  head -c 200 $O/bench_n$n.json; echo
done
