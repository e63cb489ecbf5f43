#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/gpus.txt; cat gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -8 | cut -c1-400 > gpurun_out/pytest_mgpu.log; cat gpurun_out/pytest_mgpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py 2>&1 | grep -v "^W\|^\*" | tail -8 | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -c 1500 gpurun_out/bench2.json; tail -3 gpurun_out/bench2.err | cut -c1-300
