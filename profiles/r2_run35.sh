#!/bin/bash
S=$(date +%s); timeout 900 python bench.py --batch 148 > gpurun_out/r2_bench_b148.json 2> gpurun_out/r2_bench_b148.err; echo "bench wall $(( $(date +%s) - S )) s"; tail -c 300 gpurun_out/r2_bench_b148.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_b148.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
print('e2e', {k:d['e2e'][k] for k in ('value','h2d_gbs','wall_s')}); print('e2e_path', d.get('e2e_path_inputs'))
print('fps', d['ops']['per_op_ms_per_step']['fps'], 'cpu', d['cpu_baseline'])
print(d['roofline']['families'])
PY
nvidia-smi --query-gpu=memory.used --format=csv | tail -1
