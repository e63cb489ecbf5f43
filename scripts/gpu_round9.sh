#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_full_gpu.py -x -q -m gpu 2>&1 | tail -12 | cut -c1-300 > gpurun_out/pytest_full.log; cat gpurun_out/pytest_full.log
rm -f gpurun_out/sweep_bulk.jsonl
timeout 900 python scripts/sweep_full.py --bricks "4,4,4;7,7,7" --chunks "2,4,8,16,32" --variants "18" --steps 5 --out gpurun_out/sweep_bulk.jsonl > gpurun_out/sweep_bulk.log 2>&1; cut -c1-260 gpurun_out/sweep_bulk.jsonl
timeout 600 python scripts/sweep_full.py --bricks "4,4,4" --chunks "32" --variants "2" --steps 5 --out gpurun_out/sweep_bulk.jsonl >> gpurun_out/sweep_bulk.log 2>&1; tail -1 gpurun_out/sweep_bulk.jsonl | cut -c1-260
timeout 600 python scripts/tucker_bench.py --steps 4 > gpurun_out/tucker_bench.jsonl 2>&1; cat gpurun_out/tucker_bench.jsonl | cut -c1-400
timeout 900 python -m pytest tests/test_tucker_gpu.py -x -q -m gpu 2>&1 | tail -5 | cut -c1-300
