#!/bin/bash
python bench.py --workload dsec --no-cpu-baseline --no-e2e --no-model --no-ref-cuda > gpurun_out/r2_bench_dsec2.json 2> gpurun_out/r2_bench_dsec2.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2_bench_dsec2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['ops']['per_op_ms_per_step'])
P
