#!/bin/bash
# usage: r2_scale.sh N   — weak and strong scaling lines at N GPUs (driver's launch line), things workload
N=$1
mkdir -p gpurun_out
for sc in weak strong; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --scaling $sc --no-cpu-baseline --no-model --no-ref-cuda > gpurun_out/r2_scale_${sc}_n$N.json 2> gpurun_out/r2_scale_${sc}_n$N.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r2_scale_${sc}_n$N.json').read().strip().splitlines()[-1])
print('$sc', d['n_gpus'], d['value'], d['ms_per_step'], d['config'].get('frame_pairs_per_gpu_per_step'), d['e2e'] and d['e2e']['value'])
P
done
