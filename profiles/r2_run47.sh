#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
S=$(date +%s); timeout 900 python bench.py > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; echo "bench wall $(( $(date +%s) - S )) s"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench9.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['config']['frame_pairs_per_gpu_per_step'])
print('e2e', d['e2e']['value'], 'path', d['e2e_path_inputs']['value'], 'model', d['model_e2e']['value'], d['model_e2e']['speedup_vs_ref_cuda'])
print(json.dumps(d['ops']['per_op_ms_per_step']))
print(d['roofline']['families'])
print(d['ops']['pointconv']['ms_per_step'], d['cpu_baseline']['value'])
PY
timeout 600 python bench.py --workload dsec --no-model --no-e2e --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_bench9_dsec.json 2>/dev/null
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench9_dsec.json') if l.startswith('{')][-1])
print('dsec', round(d['value'],1), round(d['ms_per_step'],3), d['config']['frame_pairs_per_gpu_per_step'])
PY
