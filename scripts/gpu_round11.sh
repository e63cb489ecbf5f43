#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_api_gpu.py -x -q -m gpu 2>&1 | tail -12 | cut -c1-400 > gpurun_out/pytest_host.log; cat gpurun_out/pytest_host.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_full_step_bulk -s 2 -c 1 -f -o gpurun_out/r1_k_full_step_bulk python bench.py --steps 1 --warmup 3 --no-cpu-baseline --variant 18 > gpurun_out/r1_ncu_bulk.log 2>&1; tail -2 gpurun_out/r1_ncu_bulk.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_full_step_bulk -s 2 -c 1 -f -o gpurun_out/r1_k_full_step_bulk_c8 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --variant 18 --chunk-planes 8 > gpurun_out/r1_ncu_bulk8.log 2>&1; tail -2 gpurun_out/r1_ncu_bulk8.log | cut -c1-200
