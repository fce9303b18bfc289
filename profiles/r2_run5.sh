#!/bin/bash
B200_CORR3D_CFG=4,1,4,1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -k "corr3d" 2>&1 | tail -5
echo "=== thin"; B200_CORR3D_CFG=4,1,4,1 timeout 120 python profiles/microbench/corr3d_time.py 2>&1 | tail -6
bash profiles/r2_run7.sh
