#!/bin/bash
# ncu evidence for profiles/: launch list of one bench solve + full captures of the two Krylov-step kernels.
set -e
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 445 -c 445 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gram_row -s 95 -c 3 -o gpurun_out/prof_gram $CMD > gpurun_out/ncu_gram.log 2>&1
$CMD > gpurun_out/ncu_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lanczos_ttr -s 95 -c 2 -o gpurun_out/prof_ttr $CMD > gpurun_out/ncu_ttr.log 2>&1
ls -la gpurun_out/
