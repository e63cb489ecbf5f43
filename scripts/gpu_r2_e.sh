#!/bin/bash
# whole GPU suite, default bench, launch list, ncu of the Tucker kernel (both Gram variants) and of the PCG kernel
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -25 | cut -c1-700 > gpurun_out/r2_pytest_all.log; cat gpurun_out/r2_pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 1200 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_ref.json 2>&1; tail -c 600 gpurun_out/r2_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_default_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
grep -c k_full_step gpurun_out/r2_launches_default_bench.csv
for g in dmma dfma; do
  VT_TUCKER_GRAM=$g timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tucker -s 3 -c 1 -f -o gpurun_out/r2_k_tucker_32_$g python scripts/tucker_bench.py --steps 2 --case 1 > gpurun_out/r2_ncu_tucker_$g.log 2>&1; tail -1 gpurun_out/r2_ncu_tucker_$g.log | cut -c1-200
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pcg -s 2 -c 1 -f -o gpurun_out/r2_k_pcg python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-tucker > gpurun_out/r2_ncu_pcg.log 2>&1; tail -1 gpurun_out/r2_ncu_pcg.log | cut -c1-200
