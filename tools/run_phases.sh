#!/bin/bash
# Per-phase device times (TK_FLAG_TIME_ALL) for one rank's share at N = 8 (128 modes), nmax 64 and 256; the per-mode
# variant at C5; and the ncu launch list of the 128-mode solve (what one rank of the 8-GPU run executes).
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
: > $O/phases.jsonl
python tools/profile_phases.py 128 10000 64 >> $O/phases.jsonl 2>> $O/phases.err
python tools/profile_phases.py 128 10000 256 >> $O/phases.jsonl 2>> $O/phases.err
python tools/profile_phases.py 1024 10000 64 >> $O/phases.jsonl 2>> $O/phases.err
python tools/profile_phases.py 1024 10000 64 reorth 1 >> $O/phases.jsonl 2>> $O/phases.err
python tools/profile_phases.py 50 1000 256 >> $O/phases.jsonl 2>> $O/phases.err
python tools/profile_phases.py 100 2000 120 arnoldi >> $O/phases.jsonl 2>> $O/phases.err
cat $O/phases.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file $O/r02_launches_d128.csv python bench.py --d 128 --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_launches_d128.log 2>&1
