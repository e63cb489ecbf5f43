#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_tucker_gpu.py -x -q -m gpu 2>&1 | tail -5 | cut -c1-300
timeout 1200 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -25 | cut -c1-400 > gpurun_out/pytest_mgpu.log; cat gpurun_out/pytest_mgpu.log
timeout 600 python scripts/tucker_bench.py --steps 4 2>&1 | grep case | cut -c1-300 | tee gpurun_out/tucker_bench4.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; python -c "
import json; d=json.loads(open('gpurun_out/bench2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['clocks'])"
