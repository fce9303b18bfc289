#!/bin/bash
for M in 4 5; do
  B200_MAX_STREAMS=$M timeout 600 python bench.py --steps 10 --warmup 3 --no-model --no-e2e --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_ms_$M.json 2> gpurun_out/r2_ms_$M.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_ms_$M.json') if l.startswith('{')][-1])
print('max_streams', $M, round(d['value'],1), 'fp/s', round(d['ms_per_step'],3), 'ms/step')
PY
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
S=$(date +%s); timeout 900 python bench.py > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; echo "bench wall $(( $(date +%s) - S )) s"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench8.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['config']['frame_pairs_per_gpu_per_step'])
print('e2e', d['e2e']['value'], 'path', d['e2e_path_inputs']['value'], 'model', d['model_e2e']['value'], d['model_e2e']['speedup_vs_ref_cuda'])
print(d['roofline']['families'])
PY
