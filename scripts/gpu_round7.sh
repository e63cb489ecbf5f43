#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 900 python -m pytest tests/test_full_gpu.py tests/test_poisson_gpu.py -x -q -m gpu 2>&1 | tail -3 | cut -c1-300; done > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
python scripts/debug_async.py 2>&1 | grep rel | cut -c1-200
rm -f gpurun_out/sweep_full.jsonl
timeout 900 python scripts/sweep_full.py --bricks "4,4,4;7,7,7" --chunks "4,8,32" --variants "2" --steps 5 > gpurun_out/sweep.log 2>&1; tail -8 gpurun_out/sweep.log | cut -c1-250
