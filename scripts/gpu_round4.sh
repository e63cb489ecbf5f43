#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log | cut -c1-400
rm -f gpurun_out/sweep_full.jsonl
timeout 1500 python scripts/sweep_full.py --bricks "4,4,4;7,7,7" --chunks "2,4,8,32" --variants "0,2,10" --steps 5 > gpurun_out/sweep.log 2>&1; tail -30 gpurun_out/sweep.log | cut -c1-300
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_full_step --csv --log-file gpurun_out/traffic_sweep.csv \
    python scripts/sweep_full.py --hexes 28 28 28 --bricks "4,4,4;7,7,7" --chunks "2,4,8,32" --variants "2" --steps 1 --out gpurun_out/sweep_under_ncu.jsonl > gpurun_out/traffic_sweep.log 2>&1
tail -3 gpurun_out/traffic_sweep.log | cut -c1-300
