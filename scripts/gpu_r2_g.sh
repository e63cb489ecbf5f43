#!/bin/bash
# slab-streaming Tucker kernel: parity tests, then timings with per-phase cycle counts, general kernel beside it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tucker_gpu.py -q -m gpu -x -k "large_grids or slab_kernel" 2>&1 | tail -40 | cut -c1-1500 > gpurun_out/r2g_pytest.log; cat gpurun_out/r2g_pytest.log
for k in slab; do
  VT_TUCKER_KERNEL=$k VT_TUCKER_PROFILE=1 timeout 300 python scripts/tucker_bench.py --steps 6 --case 1 > gpurun_out/r2g_timing_$k.jsonl 2> gpurun_out/r2g_phase_$k.log
  VT_TUCKER_KERNEL=$k VT_TUCKER_PROFILE=1 timeout 300 python scripts/tucker_bench.py --steps 6 --case 2 >> gpurun_out/r2g_timing_$k.jsonl 2>> gpurun_out/r2g_phase_$k.log
  cat gpurun_out/r2g_timing_$k.jsonl; tail -3 gpurun_out/r2g_phase_$k.log
done
