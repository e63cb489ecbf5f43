#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tucker_gpu.py -x -q -m gpu 2>&1 | tail -4 | cut -c1-300
timeout 600 python scripts/tucker_bench.py --steps 4 2>&1 | grep case | cut -c1-300 | tee gpurun_out/tucker_bench3.jsonl
VT_TUCKER_PROFILE=1 timeout 600 python scripts/tucker_bench.py --steps 1 2>&1 | grep -E "vt_step_tucker" | awk 'NR%2==0' | cut -c1-330 | tee gpurun_out/tucker_phases2.log
