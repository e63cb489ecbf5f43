#!/bin/bash
# snapshot round trip (lazy re-compression of batched uploads), then the slab kernel on 32^3 (256 threads, two CTAs per SM)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_host_api_gpu.py tests/test_tucker_gpu.py -q -m gpu -x -k "snapshot or roundtrip or generic_entry or errors or max_rank" 2>&1 | tail -4 | cut -c1-300
for m in 33 17; do
VT_TUCKER_SLAB_MIN_N=$m VT_TUCKER_PROFILE=1 timeout 300 python scripts/tucker_bench.py --steps 6 --case 1 2> gpurun_out/r2m_phase_$m.log; tail -1 gpurun_out/r2m_phase_$m.log | cut -c1-420
done
VT_TUCKER_SLAB_MIN_N=17 timeout 600 python -m pytest tests/test_tucker_gpu.py -q -m gpu -x -k "large_grids or slab_kernel" 2>&1 | tail -3 | cut -c1-300
