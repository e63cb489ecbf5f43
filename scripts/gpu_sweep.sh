#!/bin/bash
# Sweep of the full-format step kernel on the C4 share: brick / pencil order x planes per item.
mkdir -p gpurun_out
rm -f gpurun_out/sweep_order.jsonl
# whole-row items (the default kernel): tet order only
timeout 900 python scripts/sweep_full.py --bricks "4,4,4;2,2,2;p2,1;p4,1;p4,4;p7,1" --chunks "32" --variants "64" --steps 8 --out gpurun_out/sweep_order.jsonl > gpurun_out/sweep_order.log 2>&1
# velocity-chunked items
timeout 1200 python scripts/sweep_full.py --bricks "7,7,4;7,7,7;14,7,7;p7,4;p7,7;p14,4" --chunks "4,8" --variants "50" --steps 8 --out gpurun_out/sweep_order.jsonl >> gpurun_out/sweep_order.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/sweep_order.jsonl'):
    d=json.loads(l); print(d['brick'], d['chunk_planes'], d['variant'], round(d['kernel_ms'],2), round(d['frac'],3))
PY
