#!/bin/bash
# Round 2, call 1: parity of the reworked step kernel, then order / chunk / neighbour-load sweep, then ncu.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 1200 python -m pytest tests/test_full_gpu.py tests/test_virtual_ranks_gpu.py -q -m gpu 2>&1 | tail -15 | cut -c1-400 > gpurun_out/r2_pytest_full.log; cat gpurun_out/r2_pytest_full.log
rm -f gpurun_out/r2_sweep_k1.jsonl
timeout 1500 python scripts/sweep_full.py --bricks "4,4,4;p4,1;p4,4;p2,1;7,7,4" --chunks "32,16,8,4" --variants "64,192,320,448" --steps 6 --out gpurun_out/r2_sweep_k1.jsonl > gpurun_out/r2_sweep_k1.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_sweep_k1.jsonl'):
    d=json.loads(l); print(d['brick'], d['chunk_planes'], d['variant'], round(d['kernel_ms'],2), round(d['frac'],3))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_full_step_bulk -s 2 -c 1 -f -o gpurun_out/r2_k_full_step_nbd python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-coupled > gpurun_out/r2_ncu_nbd.log 2>&1; tail -2 gpurun_out/r2_ncu_nbd.log | cut -c1-300
