#!/bin/bash
export B200_CORR3D_CFG=4,1,4,1
LEVELS=0 BATCH=74 bash profiles/ncu_kernel.sh "corr3d|pointwise" 6 r2_corr3d_v2c python profiles/microbench/corr3d_time.py
