#!/bin/bash
mkdir -p gpurun_out
VT_VARIANT=18 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py 2>&1 | grep -v "^W\|^\*\|OMP_NUM" | tail -30 | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/mgpu_tucker_check.py 2>&1 | grep -v "^W\|^\*\|OMP_NUM" | tail -20 | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tests/mgpu_loop_check.py 2>&1 | grep -v "^W\|^\*\|OMP_NUM" | tail -20 | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; python -c "
import json; d=json.loads(open('gpurun_out/bench2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['clocks'])"
