#!/bin/bash
# batch sweep of the device-resident step with the end-of-round library
for B in 37 74 111 148 222; do
  timeout 600 python bench.py --batch $B --steps 10 --warmup 3 --no-model --no-e2e --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_sweep_$B.json 2> gpurun_out/r2_sweep_$B.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_sweep_$B.json') if l.startswith('{')][-1])
print($B, round(d['value'],1), 'fp/s', round(d['ms_per_step'],3), 'ms/step', 'fps', d['ops']['per_op_ms_per_step']['fps'])
PY
done
