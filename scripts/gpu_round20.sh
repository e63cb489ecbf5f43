#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_tucker_gpu.py -q -m gpu -k "large_grids" 2>&1 | tail -40 | cut -c1-300
