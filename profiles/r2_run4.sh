#!/bin/bash
mkdir -p gpurun_out
LEVELS=0 BATCH=74 bash profiles/ncu_kernel.sh "corr3d_v2_stage" 2 r2_corr3d_v2 python profiles/microbench/corr3d_time.py
