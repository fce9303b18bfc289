#!/bin/bash
for o in ev_corr corr_ev ev_only; do
  B200_SIDE_ORDER=$o timeout 600 python bench.py --steps 10 --warmup 3 --no-model --no-e2e --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_so_$o.json 2> gpurun_out/r2_so_$o.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_so_$o.json') if l.startswith('{')][-1])
print('$o', round(d['value'],1), 'fp/s', round(d['ms_per_step'],3), 'ms/step')
PY
done
