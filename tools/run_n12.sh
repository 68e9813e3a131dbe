#!/bin/bash
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_n1.json 2> gpurun_out/b_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/b_n2.json 2> gpurun_out/b_n2.err
for f in gpurun_out/b_n1.json gpurun_out/b_n2.json; do python - "$f" <<'PY'
import json,sys
for line in open(sys.argv[1]):
    line=line.strip()
    if line.startswith("{"):
        d=json.loads(line); print(sys.argv[1], d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), round(d["roofline"]["frac"],3), d["config"]["relres_last"])
    elif line: print("STDOUT NOISE:", line)
PY
done
