#!/bin/bash
# one ncu capture of the slab-streaming Tucker kernel with source correlation (C5-like 48^3 case)
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:slab --launch-skip 2 -c 1 -f -o gpurun_out/r2i_slab python scripts/tucker_bench.py --steps 4 --case 2 > gpurun_out/r2i_ncu.log 2>&1; tail -5 gpurun_out/r2i_ncu.log
ls -la gpurun_out/*.ncu-rep
