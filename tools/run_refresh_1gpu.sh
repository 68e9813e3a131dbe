#!/bin/bash
# One-GPU refresh after a change of the Krylov-step kernels: tests, smoke, the bench line, configs 1-3, the ncu captures
# of the two step kernels (CSV exports only) and the launch list of one warm solve.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
rm -f $O/*.ncu-rep $O/configs.jsonl
( time timeout 1800 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1
tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
head -c 300 $O/bench_n1.json; echo
for c in C1 C2 C3; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline >> $O/configs.jsonl 2>> $O/configs.err
done
timeout 300 python bench.py --weak --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_weak_n1.json 2>> $O/configs.err
ncu_csv() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o $O/$name "$@" > $O/ncu_$name.log 2>&1
  if [ -f $O/$name.ncu-rep ]; then
    ncu -i $O/$name.ncu-rep --page raw --csv > $O/$name.raw.csv 2>/dev/null
    ncu -i $O/$name.ncu-rep --page source --csv > $O/$name.source.csv 2>/dev/null
    rm -f $O/$name.ncu-rep
  fi
}
ncu_csv r02_gram 'gram_row_kernel' 20 2 python tools/ncu_case.py c5
ncu_csv r02_ttr 'lanczos_ttr_bulk_kernel' 20 2 python tools/ncu_case.py c5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file $O/r02_launches_bench_n1.csv python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_launches.log 2>&1
du -sh $O
