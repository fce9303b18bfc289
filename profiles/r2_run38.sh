#!/bin/bash
for P in 1 2 3 4; do
 for F in "" 3; do
  B200_FPS_T=$F timeout 600 python bench.py --split $P --steps 10 --warmup 3 --no-model --no-e2e --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_split_$P$F.json 2> gpurun_out/r2_split_$P$F.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_split_$P$F.json') if l.startswith('{')][-1])
    print('split', $P, 'fps_knob', '$F', round(d['value'],1), 'fp/s', round(d['ms_per_step'],3), 'ms/step')
except Exception as e:
    print('split', $P, 'failed', e); print(open('gpurun_out/r2_split_$P$F.err').read()[-800:])
PY
 done
done
