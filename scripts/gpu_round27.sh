#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_full_gpu.py -x -q -m gpu 2>&1 | tail -3 | cut -c1-300
for i in 1 2; do timeout 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench1.json 2>gpurun_out/bench1.err; python -c "
import json; d=json.loads(open('gpurun_out/bench1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches'], d['clocks'])"; done
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_full_step_bulk -s 2 -c 1 --csv --log-file gpurun_out/traffic_stcs.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; tail -4 gpurun_out/traffic_stcs.csv | cut -d, -f13-15
