#!/bin/bash
mkdir -p gpurun_out
VT_STEP_MIXED=1 VT_VARIANT=18 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py 2>&1 | grep "mgpu_check\|MGPU" | cut -c1-200
for m in 0 1; do VT_STEP_MIXED=$m timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench2_m$m.json 2> gpurun_out/bench2.err; python -c "
import json; d=json.loads(open('gpurun_out/bench2_m$m.json').read().strip().splitlines()[-1]); print('mixed=$m', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches'], d['clocks'])"; done
