mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_d.log
tail -15 gpurun_out/tests_d.log
timeout 600 python tools/config_report.py > gpurun_out/config_report_d.jsonl 2> gpurun_out/config_report_d.err; echo "report rc=$?"
cat gpurun_out/config_report_d.jsonl | cut -c1-400
