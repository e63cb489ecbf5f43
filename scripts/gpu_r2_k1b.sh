#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_virtual_ranks_gpu.py tests/test_poisson_gpu.py -q -m gpu 2>&1 | tail -8 | cut -c1-400 > gpurun_out/r2_pytest_full.log; cat gpurun_out/r2_pytest_full.log
rm -f gpurun_out/r2_sweep_k1b.jsonl
timeout 900 python scripts/sweep_full.py --bricks "4,4,4;p4,1" --chunks "32,8" --variants "64,192,50,18" --steps 6 --out gpurun_out/r2_sweep_k1b.jsonl > gpurun_out/r2_sweep_k1b.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_sweep_k1b.jsonl'):
    d=json.loads(l); print(d['brick'], d['chunk_planes'], d['variant'], round(d['kernel_ms'],2), round(d['frac'],3))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_full_step_bulk -s 2 -c 1 -f -o gpurun_out/r2_k_full_step_nbd_regs python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-coupled > gpurun_out/r2_ncu_nbd_regs.log 2>&1; tail -2 gpurun_out/r2_ncu_nbd_regs.log | cut -c1-300
