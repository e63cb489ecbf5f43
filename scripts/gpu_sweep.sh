#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/sweep_bulk5.jsonl
timeout 1200 python scripts/sweep_full.py --bricks "4,4,4;7,4,4;7,7,4;7,7,7;14,7,7" --chunks "2,4,8,16" --variants "50" --steps 8 --out gpurun_out/sweep_bulk5.jsonl > gpurun_out/sweep_bulk5.log 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/sweep_bulk5.jsonl'):
    d=json.loads(l); print(d['brick'], d['chunk_planes'], round(d['kernel_ms'],2), round(d['frac'],3))
PY
