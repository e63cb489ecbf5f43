#!/bin/bash
# round-2 weak scaling point at 8 GPUs (the full 1,053,696-tet C4 mesh) with the coupled loop and the Tucker leg,
# plus the strong-scaling point of the same mesh; driver-style launch
mkdir -p gpurun_out
nvidia-smi -L | wc -l
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench$N.json 2> gpurun_out/r2_bench$N.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench$N.json').read().strip().splitlines()[-1]); print($N, d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['coupled_loop']['ms_per_iteration'], d['coupled_loop']['poisson_ms'], d['tucker']['value'], d['tucker']['ms_per_step'])" || tail -5 gpurun_out/r2_bench$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-tucker --no-cpu-baseline > gpurun_out/r2_bench${N}_strong.json 2> gpurun_out/r2_bench${N}_strong.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench${N}_strong.json').read().strip().splitlines()[-1]); print('strong', $N, d['value'], d['ms_per_step'], d.get('scaling'))" || tail -5 gpurun_out/r2_bench${N}_strong.err
