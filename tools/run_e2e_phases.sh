#!/bin/bash
# GPU call 17 (1 GPU): after the NVTX ranges: tests; where the end-to-end call spends its time (per phase, 1024 and 128
# modes); the CPU arm in the reference's own complexity (flavour A) for configs 1, 2 and 4.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
( time timeout 1800 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1
tail -3 $O/pytest_gpu.log
python tools/profile_e2e.py 1024 > $O/e2e_phases_d1024.log 2>&1
python tools/profile_e2e.py 128 > $O/e2e_phases_d128.log 2>&1
tail -2 $O/e2e_phases_d128.log
: > $O/cpu_flavour_a.jsonl
for c in C1 C2 C4; do
  timeout 400 python bench.py --impl reference --config $c --cpu-flavour A --steps 1 --warmup 0 >> $O/cpu_flavour_a.jsonl 2>> $O/cpu_flavour_a.err
done
cut -c1-200 $O/cpu_flavour_a.jsonl
