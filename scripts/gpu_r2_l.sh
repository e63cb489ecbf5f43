#!/bin/bash
# outflow-half halo push: virtual-rank parity (bit identity with one context), full-format tests, Tucker sweep
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_virtual_ranks_gpu.py tests/test_multigpu_gpu.py tests/test_tucker_gpu.py -q -m gpu -x 2>&1 | tail -6 | cut -c1-300 > gpurun_out/r2l_pytest.log; cat gpurun_out/r2l_pytest.log
timeout 600 python bench.py --format tucker --tucker-sweep --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2l_tucker_sweep.json 2> gpurun_out/r2l_tucker_sweep.err
python -c "
import json; d=json.loads(open('gpurun_out/r2l_tucker_sweep.json').read().strip().splitlines()[-1])
for p in d['sweep']: print(p)"
