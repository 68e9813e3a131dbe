#!/bin/bash
# 8-GPU call: strong scaling of config 5 at N = 8 and 4, config 3 on 8 GPUs, nmax = 256 on 8 GPUs.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
P=29617
run() {  # ngpus outfile args...
  local n=$1 out=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n "$@" > $O/$out.json 2> $O/$out.err
  P=$((P+1))
  head -c 400 $O/$out.json; echo
}
run 8 bench_n8 --steps 20 --warmup 3 --no-cpu-baseline
run 4 bench_n4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras
run 8 bench_c3_n8 --config C3 --steps 20 --warmup 3 --no-cpu-baseline --no-extras
run 8 bench_n8_nmax256 --nmax 256 --steps 3 --warmup 3 --no-cpu-baseline --no-extras
TK_PEER=0 run 8 bench_n8_nccl --steps 20 --warmup 3 --no-cpu-baseline --no-extras
tail -2 $O/bench_n8.err
