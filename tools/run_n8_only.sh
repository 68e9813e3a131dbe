#!/bin/bash
# 8-GPU call, one line: config 5 at N = 8 (device-resident value and the end-to-end arm).
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29817 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_n8.json 2> $O/bench_n8.err
head -c 300 $O/bench_n8.json; echo
