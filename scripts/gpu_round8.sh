#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tucker_gpu.py -x -q -m gpu 2>&1 | tail -25 | cut -c1-400 > gpurun_out/pytest_tucker.log; cat gpurun_out/pytest_tucker.log
timeout 600 python scripts/tucker_bench.py --steps 4 > gpurun_out/tucker_bench.jsonl 2>&1; cat gpurun_out/tucker_bench.jsonl | cut -c1-400
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_tucker_gpu.py 2>&1 | tail -8 | cut -c1-300 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
# the reference's own poisson_test compiled against the host classes
mkdir -p /tmp/pt/run /tmp/pt/data/meshes/poisson_tests && cp tests/data/box_4955_tets.msh tests/data/sphere_2697_tets.msh /tmp/pt/data/meshes/poisson_tests/
(cd /tmp/pt/run && timeout 300 stdbuf -oL -eL /root/repo/vlasovtucker_b200/build/poisson_test 2>&1 | grep -E "Mesh:|MSE|Start|what|terminate|rror" | head -20) > gpurun_out/poisson_test.log 2>&1; cat gpurun_out/poisson_test.log
# launch list of the default bench command + one full capture of the dominant kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1_bench_under_ncu.log 2>&1
grep -c k_full_step gpurun_out/r1_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_full_step -s 2 -c 1 -f -o gpurun_out/r1_k_full_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_full.log 2>&1; tail -3 gpurun_out/r1_ncu_full.log
