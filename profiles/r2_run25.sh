#!/bin/bash
# PointConv second generation: parity first, then timing (v2 vs v1)
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k pointconv 2>&1 | tail -15
echo "== v2"; timeout 300 python profiles/microbench/pointconv_time.py 2>&1 | grep -v Warn
echo "== v1"; B200_POINTCONV_V1=1 timeout 300 python profiles/microbench/pointconv_time.py 2>&1 | grep -v Warn | tail -1
