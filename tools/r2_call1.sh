#!/bin/bash
# GPU call 1 of round 2: tests, bench, A/B of the new enqueue paths, ncu of the secondary kernels.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
tail -c 3000 $O/bench_n1.json
for cfg in "TK_GRAPH=0" "TK_GRAM_BALANCED=0" "TK_GRAPH=0 TK_GRAM_BALANCED=0" "TK_SEG=64" "TK_GRAM_PER_SM=1"; do
  echo "== $cfg" >> $O/ab.log
  env $cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline >> $O/ab.log 2>> $O/ab.err
done
for c in C1 C2 C3 C4; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-extras --no-cpu-baseline >> $O/configs.jsonl 2>> $O/configs.err
done
# ncu: secondary kernels at C5 and C4 sizes (one --set full pass each, few launches)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'assemble_cp|tridiag_eig_bisect|combine_chunk' -s 54 -c 6 -f -o $O/r02_c5_side python tools/ncu_case.py c5 > $O/ncu_c5_side.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'basis_mul_all' -c 2 -f -o $O/r02_c5_basis_mul python tools/ncu_case.py c5 > $O/ncu_c5_bm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'arnoldi_mgs|expm_fused|gram_blocks|expm_apply' -s 120 -c 8 -f -o $O/r02_c4_arnoldi python tools/ncu_case.py c4 > $O/ncu_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gram_row_balanced|lanczos_ttr_bulk' -s 40 -c 4 -f -o $O/r02_c5_krylov python tools/ncu_case.py c5 > $O/ncu_c5_krylov.log 2>&1
ls -la $O
