#!/bin/bash
# launch list of ONE bench step (eager: ncu cannot follow a stream capture), round 2
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-model --no-ref-cuda --eager > gpurun_out/r2_bench_under_ncu.log 2>&1 || true
python - <<'P'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2_launches.csv')) if len(r) > 14 and r[0].isdigit()]
print(len(rows), 'launches in the capture')
P
