#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "knn" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-model --no-e2e --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_ct.json 2> gpurun_out/r2_ct.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_ct.json') if l.startswith('{')][-1])
print(round(d['value'],1), 'fp/s', round(d['ms_per_step'],3), 'ms/step')
print({k:v for k,v in d['ops']['per_op_ms_per_step'].items() if k.startswith('knn')})
PY
