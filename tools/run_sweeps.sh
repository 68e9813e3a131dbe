#!/bin/bash
# Sweeps of BASELINE.json's config 5 on one GPU (d = 1024, n = 10^4): Krylov subspace size nmax and exp-sum rank t.
# Writes one JSON line per point to gpurun_out/r02_c5_sweep.jsonl; tools/summarize_sweep.py turns it into
# profiles/r02_c5_sweep.json.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
: > $O/r02_c5_sweep.jsonl
for nmax in 64 128 256; do
  timeout 600 python bench.py --nmax $nmax --steps 3 --warmup 3 --no-extras --no-cpu-baseline >> $O/r02_c5_sweep.jsonl 2>> $O/sweep.err
done
for t in 1 2 4 8 12 16 24 32 48 63; do
  timeout 600 python bench.py --t-override $t --steps 3 --warmup 3 --no-extras --no-cpu-baseline >> $O/r02_c5_sweep.jsonl 2>> $O/sweep.err
done
timeout 600 python bench.py --variant lanczos --steps 5 --warmup 3 --no-extras --no-cpu-baseline >> $O/r02_c5_sweep.jsonl 2>> $O/sweep.err
timeout 600 python bench.py --per-mode --steps 3 --warmup 3 --no-extras --no-cpu-baseline >> $O/r02_c5_sweep.jsonl 2>> $O/sweep.err
wc -l $O/r02_c5_sweep.jsonl
