set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_a.log
tail -5 gpurun_out/tests_a.log
timeout 600 bash tools/tune_ttr_bulk.sh > gpurun_out/tune_ttr_bulk.log 2>&1
cat gpurun_out/tune_ttr_bulk.log
timeout 300 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
cat gpurun_out/bench_a.json
