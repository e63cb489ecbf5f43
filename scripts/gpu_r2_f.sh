#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tucker_gpu.py tests/test_host_api_gpu.py -q -m gpu --timeout 600 -k "tiny or small_compression or group or sheath_driver_tucker" 2>&1 | tail -25 | cut -c1-700 > gpurun_out/r2_pytest_f.log; cat gpurun_out/r2_pytest_f.log
timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 300 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_ref.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_default_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
grep -c k_full_step gpurun_out/r2_launches_default_bench.csv
for g in dmma dfma; do
  VT_TUCKER_GRAM=$g timeout 600 ncu --set full --clock-control none -k regex:k_tucker -s 3 -c 1 -f -o /tmp/k_tucker_32_$g python scripts/tucker_bench.py --steps 2 --case 1 > gpurun_out/r2_ncu_tucker_$g.log 2>&1
  python scripts/ncu_summary.py /tmp/k_tucker_32_$g.ncu-rep gpurun_out/r2_k_tucker_32_${g}_summary.json > /dev/null 2>&1; ls -la gpurun_out/r2_k_tucker_32_${g}_summary.json
done
timeout 600 ncu --set full --clock-control none -k regex:k_pcg -s 2 -c 1 -f -o /tmp/k_pcg python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-tucker > gpurun_out/r2_ncu_pcg.log 2>&1
python scripts/ncu_summary.py /tmp/k_pcg.ncu-rep gpurun_out/r2_k_pcg_summary.json > /dev/null 2>&1; ls -la gpurun_out/r2_k_pcg_summary.json
du -sh gpurun_out
