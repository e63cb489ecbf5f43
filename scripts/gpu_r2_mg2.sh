#!/bin/bash
# two GPUs, one process per GPU: CUDA-IPC halo + partitioned Poisson parity, then the bench at N=2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -4
export VT_COMM_TIMEOUT_MS=20000
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu --timeout 600 2>&1 | tail -25 | cut -c1-900 > gpurun_out/r2_pytest_mgpu.log; cat gpurun_out/r2_pytest_mgpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; tail -c 2500 gpurun_out/r2_bench2.json; tail -5 gpurun_out/r2_bench2.err | cut -c1-300
