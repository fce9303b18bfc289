#!/bin/bash
B200FLOW_LIB=$PWD/gpurun_variants/libb200flow_trace.so timeout 120 python profiles/microbench/corr3d_trace.py
LEVELS=0 BATCH=74 bash profiles/ncu_kernel.sh "corr3d_v2_stage2" 1 r2_corr3d_s2 python profiles/microbench/corr3d_time.py > /dev/null 2>&1
python profiles/ncu_stalls.py gpurun_out/r2_corr3d_s2_raw.csv; python profiles/ncu_hot.py gpurun_out/r2_corr3d_s2_source.csv corr3d_v2_stage2
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -k "corr2d" 2>&1 | tail -2
