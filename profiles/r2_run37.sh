#!/bin/bash
for L in 1 2 3 4; do
  B200_VOXEL_LANES=$L timeout 600 python bench.py --steps 10 --warmup 3 --no-model --no-e2e --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_lanes_$L.json 2> gpurun_out/r2_lanes_$L.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_lanes_$L.json') if l.startswith('{')][-1])
print('lanes', $L, round(d['value'],1), 'fp/s', round(d['ms_per_step'],3), 'ms/step')
PY
done
for L in 1 2; do
  B200_VOXEL_LANES=$L timeout 600 python bench.py --workload dsec --steps 10 --warmup 3 --no-model --no-e2e --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_lanes_dsec_$L.json 2> gpurun_out/r2_lanes_dsec_$L.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_lanes_dsec_$L.json') if l.startswith('{')][-1])
print('dsec lanes', $L, round(d['value'],1), 'fp/s', round(d['ms_per_step'],3), 'ms/step')
PY
done
