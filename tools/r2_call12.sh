#!/bin/bash
# GPU call 12 (1 GPU): final validation of the shipped library: tests, smoke, C4, rank sweep, fallback cost report,
# ncu of the blocked Arnoldi kernel.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
rm -f $O/*.ncu-rep $O/configs.jsonl
( time timeout 1800 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1
tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
for c in C4 C2; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline >> $O/configs.jsonl 2>> $O/configs.err
done
timeout 300 python bench.py --config C4 --steps 3 --warmup 2 --impl reference > $O/bench_reference_arm_c4.json 2>> $O/configs.err
bash tools/run_sweeps.sh > $O/sweeps.log 2>&1
timeout 600 python tools/fallback_report.py > $O/fallback.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:arnoldi_bgs_kernel -s 28 -c 2 -f -o $O/r02_bgs python tools/ncu_case.py c4 > $O/ncu_r02_bgs.log 2>&1
if [ -f $O/r02_bgs.ncu-rep ]; then
  ncu -i $O/r02_bgs.ncu-rep --page raw --csv > $O/r02_bgs.raw.csv 2>/dev/null
  ncu -i $O/r02_bgs.ncu-rep --page source --csv > $O/r02_bgs.source.csv 2>/dev/null
  rm -f $O/r02_bgs.ncu-rep
fi
du -sh $O
