#!/bin/bash
# 8 GPUs: two launches (boundary tets, then interior) against one mixed queue (VT_STEP_MIXED=1), full format only
mkdir -p gpurun_out
for M in 0 1; do
VT_STEP_MIXED=$M timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2953$M bench.py --gpus 8 --steps 10 --warmup 3 --no-tucker --no-coupled --no-cpu-baseline > gpurun_out/r2_bench8_mixed$M.json 2> gpurun_out/r2_bench8_mixed$M.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench8_mixed$M.json').read().strip().splitlines()[-1]); print('mixed', $M, d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])" || tail -5 gpurun_out/r2_bench8_mixed$M.err
done
