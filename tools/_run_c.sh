mkdir -p gpurun_out
for cfg in "TK_TTR_BULK=1" "TK_TTR_BULK=0" ; do
  for d in 1024 256; do
    echo "== $cfg d=$d"; env $cfg timeout 120 python tools/_repro_ttt.py $d 30 2>&1 | tail -2
  done
done
timeout 300 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench rc=$?"
cat gpurun_out/bench_c.json; tail -3 gpurun_out/bench_c.err
