#!/bin/bash
# slab-streaming Tucker kernel: parity tests, timings with per-phase cycle counts, one ncu capture with source view
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tucker_gpu.py -q -m gpu -x -k "large_grids or slab_kernel" 2>&1 | tail -40 | cut -c1-1500 > gpurun_out/r2h_pytest.log; cat gpurun_out/r2h_pytest.log
VT_TUCKER_PROFILE=1 timeout 300 python scripts/tucker_bench.py --steps 6 --case 1 > gpurun_out/r2h_timing_slab.jsonl 2> gpurun_out/r2h_phase_slab.log
VT_TUCKER_PROFILE=1 timeout 300 python scripts/tucker_bench.py --steps 6 --case 2 >> gpurun_out/r2h_timing_slab.jsonl 2>> gpurun_out/r2h_phase_slab.log
cat gpurun_out/r2h_timing_slab.jsonl; tail -2 gpurun_out/r2h_phase_slab.log
if [ -n "$NCU" ]; then
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tucker_slab --launch-skip 3 -c 1 -f -o gpurun_out/r2h_slab python scripts/tucker_bench.py --steps 4 --case 2 > gpurun_out/r2h_ncu.log 2>&1; tail -3 gpurun_out/r2h_ncu.log
fi
