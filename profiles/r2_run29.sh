#!/bin/bash
export B200FLOW_LIB=gpurun_variants/libb200flow_pc_NOCOPYNOMMA.so
bash profiles/ncu_kernel.sh pointconv_v2_kernel 1 r2_pc2_pure python profiles/microbench/pointconv_one.py 2>&1 | tail -12
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2_pc2_pure_source.csv')))
R=[r for r in rows if r and r[0].startswith('0x')]
tot=sum(int(r[2]) for r in R)
print('top stall instructions (of', tot, 'samples)')
for i,r in sorted(enumerate(R), key=lambda t:-int(t[1][2]))[:25]:
    print(i, r[1][:80], 'stall', r[2], 'exec', r[5])
PY
