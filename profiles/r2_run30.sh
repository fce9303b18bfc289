#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; tail -c 600 gpurun_out/r2_bench6.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench6.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
print('e2e', d['e2e']); print('e2e_path', d.get('e2e_path_inputs')); print('model', d.get('model_e2e'))
print(d['ops'].get('pointconv'))
PY
