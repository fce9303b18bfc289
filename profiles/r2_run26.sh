#!/bin/bash
bash profiles/ncu_kernel.sh pointconv_v2_kernel 1 r2_pc2_est1 python profiles/microbench/pointconv_one.py 2>&1 | tail -40
python profiles/ncu_hot.py gpurun_out/r2_pc2_est1_source.csv pointconv_v2_kernel 2>&1 | tail -40
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2_pc2_est1_source.csv')))
R=[r for r in rows if r and r[0].startswith('0x')]
tot=sum(int(r[2]) for r in R)
print('top stall instructions (of', tot, 'samples)')
for i,r in sorted(enumerate(R), key=lambda t:-int(t[1][2]))[:40]:
    print(i, r[1][:80], 'stall', r[2], 'exec', r[5])
PY
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2_pc2_est1_raw.csv')))
h=rows[0]; r=rows[2]
for k in h:
    if any(t in k for t in ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared','smsp__inst_executed_pipe','sm__inst_executed_pipe_tc','l1tex__lsu_writeback','smsp__average_warp','lts__t_sectors.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','sm__pipe_shared_cycles_active','l1tex__data_pipe','smsp__pcsamp_warps_issue_stalled')):
        print(k, r[h.index(k)])
PY
