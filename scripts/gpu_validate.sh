#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | cut -c1-300 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1800 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 700 gpurun_out/bench_ref.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_full_step_bulk -s 2 -c 1 -f -o gpurun_out/r1_k_full_step_default python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_default.log 2>&1; tail -2 gpurun_out/r1_ncu_default.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1_bench_under_ncu.log 2>&1
grep -c k_full_step gpurun_out/r1_launches.csv
