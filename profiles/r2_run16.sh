#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -k "corr3d" 2>&1 | tail -3
echo "== own d1"; timeout 120 python profiles/microbench/corr3d_time.py 2>&1 | tail -6
echo "== own d2"; B200_CORR3D_CFG=4,2,4,1 timeout 120 python profiles/microbench/corr3d_time.py 2>&1 | tail -6
echo "== no own"; B200_CORR3D_NO_OWN=1 timeout 120 python profiles/microbench/corr3d_time.py 2>&1 | tail -6
