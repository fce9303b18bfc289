#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python bench.py --no-cpu-baseline --no-e2e --no-model --no-ref-cuda > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2_bench5.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['ops']['per_op_ms_per_step']); print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline'].get('isolated'))
P
