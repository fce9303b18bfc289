#!/bin/bash
for M in 1 2 3 6 10; do
  B200_MAX_STREAMS=$M timeout 600 python bench.py --steps 10 --warmup 3 --no-model --no-e2e --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_ms_$M.json 2> gpurun_out/r2_ms_$M.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_ms_$M.json') if l.startswith('{')][-1])
print('max_streams', $M, round(d['value'],1), 'fp/s', round(d['ms_per_step'],3), 'ms/step')
PY
done
