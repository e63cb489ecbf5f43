#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multigpu_gpu.py tests/test_full_gpu.py -x -q -m gpu 2>&1 | tail -6 | cut -c1-400
for i in 1 2; do timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; python -c "
import json; d=json.loads(open('gpurun_out/bench2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches'], d['clocks'])"; done
for i in 1 2; do timeout 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench1.json 2>gpurun_out/bench1.err; python -c "
import json; d=json.loads(open('gpurun_out/bench1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches'], d['clocks'])"; done
