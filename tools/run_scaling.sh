#!/bin/bash
# 1/2/4/8-GPU strong-scaling sweep of bench.py (as the driver launches it)
for N in 1 2 4 8; do
  if [ "$N" -gt "${MAXGPUS:-8}" ]; then break; fi
  if [ "$N" = 1 ]; then
    python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  python - gpurun_out/scale_n$N.json <<'PY'
import json,sys
for line in open(sys.argv[1]):
    line=line.strip()
    if line.startswith("{"):
        d=json.loads(line); print(d["n_gpus"], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],3), "ttt", d.get("time_to_tol"))
    elif line: print("STDOUT NOISE:", line)
PY
done
