#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_full_gpu.py -x -q -m gpu 2>&1 | tail -12 | cut -c1-300 > gpurun_out/pytest_full.log; cat gpurun_out/pytest_full.log
rm -f gpurun_out/sweep_bulk3.jsonl
timeout 900 python scripts/sweep_full.py --bricks "4,4,4" --chunks "8,16,32" --variants "18,50" --steps 5 --out gpurun_out/sweep_bulk3.jsonl > gpurun_out/sweep_bulk3.log 2>&1; cut -c1-260 gpurun_out/sweep_bulk3.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_full_step_bulk -s 2 -c 1 -f -o gpurun_out/r1_k_full_step_bulk16 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --variant 18 > gpurun_out/r1_ncu_bulk16.log 2>&1; tail -2 gpurun_out/r1_ncu_bulk16.log | cut -c1-200
