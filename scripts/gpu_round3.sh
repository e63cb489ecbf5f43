#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu > gpurun_out/pytest_mgpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_mgpu.log
tail -30 gpurun_out/pytest_mgpu.log | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --chunk-planes 32 --brick 4 4 4 --variant 2 > gpurun_out/bench2.json 2> gpurun_out/bench2.err
tail -c 1500 gpurun_out/bench2.json; tail -5 gpurun_out/bench2.err | cut -c1-400
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --chunk-planes 32 --brick 4 4 4 --variant 2 --no-cpu-baseline > gpurun_out/bench1.json 2> gpurun_out/bench1.err
tail -c 600 gpurun_out/bench1.json
