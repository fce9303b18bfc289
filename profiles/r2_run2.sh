#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_dropin.py tests/test_gpu_parity.py tests/test_stack.py -m gpu -q --timeout 900 -k "model or project or sampler or stack or graph or bench or other" > gpurun_out/r2_pytest2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
tail -15 gpurun_out/r2_pytest2.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-model --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench2.json'))
print(d['value'], d['ms_per_step'])
print(d['roofline']['families'])
for e in d['roofline']['per_op']:
    if e['op'].startswith('project') or e['op'].startswith('grid'): print(e['op'], e['ms'], e['frac'])
PY
