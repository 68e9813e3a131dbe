#!/bin/bash
# One-GPU validation of the shipped library: tests, smoke, the bench line.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
( time timeout 1800 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
head -c 300 $O/bench_n1.json; echo; tail -3 $O/bench_n1.err
timeout 300 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c4.json 2>> $O/bench_n1.err
