#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k pointconv 2>&1 | tail -3
echo "== v2"; timeout 300 python profiles/microbench/pointconv_time.py 2>&1 | grep -v Warn
