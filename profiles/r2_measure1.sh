#!/bin/bash
# BASELINE.json configs[3] / [4] on one GPU: dsec (640x480, tri-linear voxels), hd (1920x1080, 32768 points, the reference's
# hard-coded pyramid) and hd_scaled (pyramid 16384..1024); + the k = 32 searches of configs[2]
mkdir -p gpurun_out
for w in dsec hd hd_scaled; do
  B=74; [ $w != dsec ] && B=18      # 36 clouds x 4-CTA clusters = 144 CTAs: one wave of the cluster FPS
  timeout 900 python bench.py --workload $w --batch $B --steps 10 --warmup 3 --no-model --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err
  echo "$w rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_$w.json'))
print('$w', round(d['value'],1), 'fp/s', round(d['ms_per_step'],3), 'ms/step batch', d['config']['frame_pairs_per_gpu_per_step'])
print('   families', {k:(v['ms'], v['frac']) for k,v in d['roofline']['families'].items()})
print('   ', {e['op']:(e['ms'], e.get('frac')) for e in d['roofline']['per_op'] if e['op'] in ('event_voxel','corr3d','fps','corr2d_L1','project_nn_corr_L1')})
PY
done
