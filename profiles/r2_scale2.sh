#!/bin/bash
# usage: r2_scale2.sh N   — the driver's launch line at N GPUs with the default arguments (bench batch 148 per GPU), wall time included
N=$1
mkdir -p gpurun_out
S=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_scale_b148_n$N.json 2> gpurun_out/r2_scale_b148_n$N.err
echo "rc=$? wall $(( $(date +%s) - S )) s"
python - <<P
import json
d=json.loads(open('gpurun_out/r2_scale_b148_n$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), d['config'].get('frame_pairs_per_gpu_per_step'), 'e2e', d['e2e'] and round(d['e2e']['value'],1), 'path', d.get('e2e_path_inputs') and d['e2e_path_inputs'].get('value'), d.get('verify'))
P
free -g | head -2
