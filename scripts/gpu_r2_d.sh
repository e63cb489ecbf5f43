#!/bin/bash
# partitioned Poisson on virtual ranks (short leash), then Tucker tests (DMMA Gram, small eps), then timings
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_virtual_ranks_gpu.py -q -m gpu -k "poisson or coupled" -x 2>&1 | tail -30 | cut -c1-900 > gpurun_out/r2_pytest_d1.log; cat gpurun_out/r2_pytest_d1.log
timeout 900 python -m pytest tests/test_tucker_gpu.py tests/test_poisson_gpu.py -q -m gpu 2>&1 | tail -30 | cut -c1-900 > gpurun_out/r2_pytest_d2.log; cat gpurun_out/r2_pytest_d2.log
./scripts/microbench/peaks > gpurun_out/r2_microbench_peaks.json 2>&1; cat gpurun_out/r2_microbench_peaks.json
for g in dmma dfma; do
  VT_TUCKER_GRAM=$g VT_TUCKER_PROFILE=1 timeout 300 python scripts/tucker_bench.py --steps 4 > gpurun_out/r2_tucker_timing_$g.jsonl 2> gpurun_out/r2_tucker_phase_$g.log; cat gpurun_out/r2_tucker_timing_$g.jsonl; tail -3 gpurun_out/r2_tucker_phase_$g.log
done
