#!/bin/bash
export B200_CORR3D_CFG=4,1,4,1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"corr3d|pointwise" --csv --log-file gpurun_out/r2_corr3d_times.csv python profiles/microbench/corr3d_time.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_corr3d_times.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
# per level: launches repeat 10x (3 warm + 7 timed); take the last occurrence group
seq=[(r[ki].split('(')[0].replace('void b200::','').replace('b200::',''), float(r[vi].replace(',','')), r[gi]) for r in rows[1:]]
per_call=7  # kernels per op call: prep_weights, 2 linears, 2 v2 preps, stage1, stage2
calls=[seq[i:i+per_call] for i in range(0,len(seq),per_call)]
for lvl in range(5):
    c=calls[lvl*10+9]
    print('level',lvl+1, ' '.join(f"{n[:24]}={t/1e3:.1f}us" for n,t,g in c), 'sum=%.1f'%(sum(t for _,t,_ in c)/1e3))
PY
