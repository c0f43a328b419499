#!/bin/bash
# Sweep of the pipelined CSR fill (element kernel of batch b+1 overlapped with the fill of batch b) against the sequential default.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python bench.py --no-extras --steps 30 --warmup 5 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['ms_per_step'], d['value'])"; }
run GG_PIPELINE=0
for b in 2 3 4 8; do for r in 0 1; do run GG_PIPELINE=1 GG_PIPELINE_BATCHES=$b GG_PIPELINE_ROOM=$r; done; done
