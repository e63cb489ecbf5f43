#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_full_gpu.py tests/test_tucker_gpu.py -x -q -m gpu 2>&1 | tail -12 | cut -c1-300 > gpurun_out/pytest_full.log; cat gpurun_out/pytest_full.log
rm -f gpurun_out/sweep_bulk2.jsonl
timeout 900 python scripts/sweep_full.py --bricks "4,4,4" --chunks "4,8,16,32" --variants "18" --steps 5 --out gpurun_out/sweep_bulk2.jsonl > gpurun_out/sweep_bulk2.log 2>&1; cut -c1-260 gpurun_out/sweep_bulk2.jsonl
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_full_step --csv --log-file gpurun_out/traffic_bulk.csv \
    python scripts/sweep_full.py --hexes 28 28 28 --bricks "4,4,4" --chunks "4,8,16,32" --variants "18" --steps 1 --out gpurun_out/sweep_bulk_under_ncu.jsonl > gpurun_out/traffic_bulk.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/traffic_bulk.csv')) if len(r)>10 and r[0].isdigit()]
from collections import defaultdict
d=defaultdict(dict)
for r in rows: d[r[0]][r[-3]]=float(r[-1].replace(',',''))
for k,v in d.items(): print(k, {a:round(b,3) for a,b in v.items()})
PY
timeout 600 python scripts/tucker_bench.py --steps 4 > gpurun_out/tucker_bench.jsonl 2>&1; cat gpurun_out/tucker_bench.jsonl | cut -c1-400
