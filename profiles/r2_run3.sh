#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -k "corr3d" 2>&1 | tail -30
echo "=== v2"; timeout 120 python profiles/microbench/corr3d_time.py
echo "=== v1"; B200_CORR3D_V1=1 timeout 120 python profiles/microbench/corr3d_time.py
