#!/bin/bash
timeout 1200 python -m pytest tests/test_stack.py tests/test_gpu_parity.py -m gpu -x -q -k "bench_batch or fps or feeder" 2>&1 | tail -3
S=$(date +%s); timeout 900 python bench.py > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; echo "bench wall $(( $(date +%s) - S )) s"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench7.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['config']['frame_pairs_per_gpu_per_step'])
print('e2e', d['e2e']['value'], 'path', d['e2e_path_inputs']['value'], 'model', d['model_e2e']['value'], d['model_e2e']['speedup_vs_ref_cuda'])
print(json.dumps(d['ops']['per_op_ms_per_step']))
print(d['roofline']['families']); print({k:v for k,v in d['roofline'].items() if k not in ('per_op','families')})
print(d['ops']['pointconv'])
PY
