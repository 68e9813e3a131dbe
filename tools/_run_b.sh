set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lanczos_ttr -s 95 -c 2 -o gpurun_out/prof_ttr_bulk $CMD > gpurun_out/ncu_ttr_bulk.log 2>&1
ls -la gpurun_out/
