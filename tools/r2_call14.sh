#!/bin/bash
# GPU call 14 (1 GPU): the shipped library, final single-GPU lines: tests, smoke, bench, reference arm, fallback report,
# weak-scaling base point (d = 128), launch list.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
rm -f $O/*.ncu-rep
( time timeout 1800 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1
tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
head -c 300 $O/bench_n1.json; echo
timeout 300 python bench.py --weak --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_weak_n1.json 2>> $O/bench_n1.err
timeout 900 python tools/fallback_report.py > $O/fallback.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file $O/r02_launches_bench_n1.csv python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_launches.log 2>&1
du -sh $O
