#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_full_gpu.py tests/test_poisson_gpu.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log | cut -c1-400
rm -f gpurun_out/sweep_full.jsonl
timeout 900 python scripts/sweep_full.py --bricks "4,4,4;7,7,7" --chunks "4,8,32" --variants "0,2" --steps 5 > gpurun_out/sweep.log 2>&1; tail -30 gpurun_out/sweep.log | cut -c1-300
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_full_step -s 2 -c 1 --csv --log-file gpurun_out/inst_async.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --hexes 14 14 14 --chunk-planes 32 --brick 7 7 7 --variant 2 > /dev/null 2>&1
cat gpurun_out/inst_async.csv | tail -6 | cut -c1-300
