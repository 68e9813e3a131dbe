#!/bin/bash
# GPU call 3 of round 2 (1 GPU): tests, bench, per-rank emulation of N=8 (d=128), C4 with the fixed blocked Arnoldi
# kernel + its ncu capture, the config-5 sweeps.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
rm -f $O/*.ncu-rep $O/ab.log $O/configs.jsonl $O/d128.log
( time timeout 1800 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1
tail -8 $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
head -c 600 $O/bench_n1.json; echo
for c in C1 C2 C3 C4; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline >> $O/configs.jsonl 2>> $O/configs.err
done
# what one rank of an 8-GPU run does, alone on a GPU (128 modes): graph replay vs stream launches, old vs balanced Gram grid
for cfg in "TK_GRAPH=1" "TK_GRAPH=0" "TK_GRAM_BALANCED=0" "TK_SEG=64"; do
  echo "== $cfg" >> $O/d128.log
  env $cfg timeout 300 python bench.py --d 128 --steps 20 --warmup 3 --no-extras --no-cpu-baseline >> $O/d128.log 2>> $O/d128.err
done
echo "== d=32 (C3 per rank at 8 GPUs)" >> $O/d128.log
timeout 300 python bench.py --d 32 --steps 20 --warmup 3 --no-extras --no-cpu-baseline >> $O/d128.log 2>> $O/d128.err
# ncu: the blocked Arnoldi kernel at C4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:arnoldi_bgs_kernel -s 28 -c 3 -f -o $O/r02_c4_bgs python tools/ncu_case.py c4 > $O/ncu_r02_c4_bgs.log 2>&1
if [ -f $O/r02_c4_bgs.ncu-rep ]; then
  ncu -i $O/r02_c4_bgs.ncu-rep --page raw --csv > $O/r02_c4_bgs.raw.csv 2>/dev/null
  ncu -i $O/r02_c4_bgs.ncu-rep --page source --csv > $O/r02_c4_bgs.source.csv 2>/dev/null
  rm -f $O/r02_c4_bgs.ncu-rep
fi
bash tools/run_sweeps.sh > $O/sweeps.log 2>&1
du -sh $O
