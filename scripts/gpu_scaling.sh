#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench$N.json 2> gpurun_out/bench$N.err; python -c "
import json; d=json.loads(open('gpurun_out/bench$N.json').read().strip().splitlines()[-1]); print($N, d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches'], d['clocks'])" || tail -5 gpurun_out/bench$N.err
done
timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -4 | cut -c1-300
