#!/bin/bash
# round-2 closing run: whole GPU suite, bench (both arms), ncu launch list of the default command
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2j_gpu.txt
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -15 | cut -c1-400 > gpurun_out/r2j_pytest_gpu.log; cat gpurun_out/r2j_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -c 400 gpurun_out/r2j_bench.json; tail -3 gpurun_out/r2j_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2j_bench_ref.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2j_launches_default_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2j_bench_under_ncu.log 2>&1
grep -c "k_full_step\|k_tucker" gpurun_out/r2j_launches_default_bench.csv
python __graft_entry__.py smoke 2>&1 | tail -2
VT_TUCKER_PROFILE=1 timeout 300 python scripts/tucker_bench.py --steps 6 > gpurun_out/r2j_tucker_timing.jsonl 2> gpurun_out/r2j_tucker_phase.log; cat gpurun_out/r2j_tucker_timing.jsonl
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:slab --launch-skip 2 -c 1 -f -o gpurun_out/r2j_slab python scripts/tucker_bench.py --steps 4 --case 2 > gpurun_out/r2j_ncu.log 2>&1; tail -2 gpurun_out/r2j_ncu.log
