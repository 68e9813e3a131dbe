#!/bin/bash
# GPU call of round 2: tests, bench, A/B of the enqueue paths, configs, ncu of the secondary kernels.
# Everything that comes back must fit in 64 MiB: ncu reports are exported to CSV on the box and deleted.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
rm -f $O/*.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
nproc > $O/nproc.txt
( time timeout 1800 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1
tail -15 $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
head -c 1500 $O/bench_n1.json; echo
rm -f $O/ab.log $O/configs.jsonl
for cfg in "TK_GRAPH=0" "TK_GRAM_BALANCED=0" "TK_SEG=64" "TK_GRAM_PER_SM=1"; do
  echo "== $cfg" >> $O/ab.log
  env $cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline >> $O/ab.log 2>> $O/ab.err
done
for c in C1 C2 C3 C4; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline >> $O/configs.jsonl 2>> $O/configs.err
done
echo "== C4 strict MGS" >> $O/ab.log
TK_MGS_BLOCK=0 timeout 300 python bench.py --config C4 --steps 5 --warmup 3 --no-extras --no-cpu-baseline >> $O/ab.log 2>> $O/ab.err
ncu_csv() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o $O/$name "$@" > $O/ncu_$name.log 2>&1
  if [ -f $O/$name.ncu-rep ]; then
    ncu -i $O/$name.ncu-rep --page raw --csv > $O/$name.raw.csv 2>/dev/null
    ncu -i $O/$name.ncu-rep --page details --csv > $O/$name.details.csv 2>/dev/null
    if [ "$5" = "keepsrc" ]; then :; fi
  fi
}
ncu_csv r02_c5_side 'assemble_cp|tridiag_eig_bisect|combine_chunk' 54 6 python tools/ncu_case.py c5
rm -f $O/r02_c5_side.ncu-rep
ncu_csv r02_c5_basis_mul 'basis_mul_all' 0 2 python tools/ncu_case.py c5
rm -f $O/r02_c5_basis_mul.ncu-rep
ncu_csv r02_c4_arnoldi 'arnoldi_|expm_fused|gram_blocks|expm_apply' 120 8 python tools/ncu_case.py c4
rm -f $O/r02_c4_arnoldi.ncu-rep
ncu_csv r02_nonsym200 'expm_fused|gram_blocks|combine_chunk|expm_apply' 560 4 python tools/ncu_case.py nonsym200
rm -f $O/r02_nonsym200.ncu-rep
ncu_csv r02_c5_krylov 'gram_row_balanced|lanczos_ttr_bulk' 40 4 python tools/ncu_case.py c5
ncu -i $O/r02_c5_krylov.ncu-rep --page source --csv -k regex:lanczos_ttr_bulk > $O/r02_c5_ttr.source.csv 2>/dev/null
rm -f $O/r02_c5_krylov.ncu-rep
# launch list of one bench solve (graph replay): shares per kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02_launches_bench_n1.csv python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_launches.log 2>&1
du -sh $O; ls -la $O
