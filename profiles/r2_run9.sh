#!/bin/bash
timeout 600 python bench.py --steps 10 --warmup 3 --no-model --no-cpu-baseline --no-ref-cuda --no-e2e > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench3.json'))
print(d['value'], d['ms_per_step'], d['roofline']['serial_ms_all_ops'])
for k,v in sorted(d['ops']['per_op_ms_per_step'].items()): print(f"  {k:24s} {v}")
PY
