#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tucker -s 4 -c 1 -f -o gpurun_out/r1_k_tucker_32 python scripts/tucker_bench.py --steps 2 --case 1 > gpurun_out/r1_ncu_tucker.log 2>&1; tail -3 gpurun_out/r1_ncu_tucker.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tucker -s 4 -c 1 -f -o gpurun_out/r1_k_tucker_11 python scripts/tucker_bench.py --steps 2 --case 0 > gpurun_out/r1_ncu_tucker11.log 2>&1; tail -3 gpurun_out/r1_ncu_tucker11.log | cut -c1-300
