#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_tucker_gpu.py tests/test_host_api_gpu.py -x -q -m gpu 2>&1 | tail -25 | cut -c1-400 > gpurun_out/pytest_tucker.log; cat gpurun_out/pytest_tucker.log
timeout 600 python scripts/tucker_bench.py --steps 4 > gpurun_out/tucker_bench2.jsonl 2>&1; cat gpurun_out/tucker_bench2.jsonl | cut -c1-400
