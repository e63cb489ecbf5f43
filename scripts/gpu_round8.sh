#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | cut -c1-300 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
# the reference's own poisson_test (needs ../data/meshes/poisson_tests): run from a scratch dir with the fixtures we carry
mkdir -p /tmp/pt/run /tmp/pt/data/meshes/poisson_tests && cp tests/data/box_4955_tets.msh tests/data/sphere_2697_tets.msh /tmp/pt/data/meshes/poisson_tests/
(cd /tmp/pt/run && timeout 300 /root/repo/vlasovtucker_b200/build/poisson_test 2>&1 | grep -E "Mesh:|MSE|Start|what|terminate" | head -20) > gpurun_out/poisson_test.log 2>&1; cat gpurun_out/poisson_test.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 600 gpurun_out/bench_ref.json
