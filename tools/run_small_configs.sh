#!/bin/bash
# Tests + the small configurations after a change to the side streams (8 eigensolver streams, history row off the
# Krylov stream): C1-C4 contract lines, TensorLanczos at C5.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
rm -f $O/configs.jsonl
( time timeout 1800 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1
tail -4 $O/pytest_gpu.log
for c in C1 C2 C3 C4; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline >> $O/configs.jsonl 2>> $O/configs.err
done
timeout 300 python bench.py --variant lanczos --steps 5 --warmup 3 --no-extras --no-cpu-baseline >> $O/configs.jsonl 2>> $O/configs.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline >> $O/configs.jsonl 2>> $O/configs.err
