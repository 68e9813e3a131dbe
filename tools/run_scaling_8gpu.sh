#!/bin/bash
# 8-GPU call (final library): strong scaling N = 8 / 4, nmax = 256 and config 3 on 8 GPUs, weak scaling N = 2 / 4.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
P=29717
run() {  # ngpus outfile args...
  local n=$1 out=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n "$@" > $O/$out.json 2> $O/$out.err
  P=$((P+1))
  head -c 260 $O/$out.json; echo
}
run 8 bench_n8 --steps 20 --warmup 3 --no-cpu-baseline
run 4 bench_n4 --steps 10 --warmup 3 --no-cpu-baseline
run 8 bench_n8_nmax256 --nmax 256 --steps 3 --warmup 3 --no-cpu-baseline --no-extras
run 8 bench_c3_n8 --config C3 --steps 20 --warmup 3 --no-cpu-baseline --no-extras
run 2 bench_weak_n2 --weak --steps 10 --warmup 3 --no-cpu-baseline --no-extras
run 4 bench_weak_n4 --weak --steps 10 --warmup 3 --no-cpu-baseline --no-extras
tail -2 $O/bench_n8.err
