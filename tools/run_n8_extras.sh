#!/bin/bash
# 8-GPU box: the secondary N = 8 lines with the shipped kernels: nmax = 256, config 3, and config 5 through the NCCL fallback.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
P=29937
run() {  # outfile args...
  local out=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 "$@" > $O/$out.json 2> $O/$out.err
  P=$((P+1))
  head -c 160 $O/$out.json; echo
}
run bench_n8_nmax256 --nmax 256 --steps 3 --warmup 3 --no-cpu-baseline --no-extras
run bench_c3_n8 --config C3 --steps 20 --warmup 3 --no-cpu-baseline --no-extras
TK_PEER=0 run bench_n8_nccl --steps 10 --warmup 3 --no-cpu-baseline --no-extras
