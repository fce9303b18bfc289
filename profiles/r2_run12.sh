#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log; tail -4 gpurun_out/r2_pytest3.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench4.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'], d['model_e2e'].get('b200'))
print(d['roofline']['families'])
PY
