#!/bin/bash
# partitioned Poisson on virtual ranks, the single-context Poisson tests, the host API tests; then the bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_virtual_ranks_gpu.py tests/test_poisson_gpu.py tests/test_host_api_gpu.py -q -m gpu 2>&1 | tail -30 | cut -c1-600 > gpurun_out/r2_pytest_c.log; cat gpurun_out/r2_pytest_c.log
timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 6000 gpurun_out/r2_bench.json; tail -5 gpurun_out/r2_bench.err
