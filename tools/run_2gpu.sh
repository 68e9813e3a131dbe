#!/bin/bash
# 2-GPU call: the multi-GPU parity tests (peer exchange vs NCCL, graph replay) and the N=2 bench line.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
( time timeout 1500 python -m pytest tests/test_gpu_multi.py -q -x ) > $O/pytest_multi.log 2>&1
tail -15 $O/pytest_multi.log
P=29517
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err
head -c 1200 $O/bench_n2.json; echo
TK_PEER=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_n2_nccl.json 2> $O/bench_n2_nccl.err
head -c 600 $O/bench_n2_nccl.json; echo
tail -3 $O/bench_n2.err
