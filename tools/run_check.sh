#!/bin/bash
# one-GPU acceptance run: parity tests, the bench line (with the CPU baseline), the five-configuration report and the
# 107-solve corpus report against the stored Julia histories
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
tail -4 gpurun_out/tests.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_ref.json
timeout 600 python tools/config_report.py > gpurun_out/config_report.jsonl 2> gpurun_out/config_report.err; echo "report rc=$?"
cut -c1-330 gpurun_out/config_report.jsonl
timeout 300 python tools/corpus_gpu_report.py --kmax 40 > gpurun_out/corpus_gpu_report.log 2>&1; echo "corpus rc=$?"
tail -2 gpurun_out/corpus_gpu_report.log
