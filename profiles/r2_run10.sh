#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -k "project or sampler" 2>&1 | tail -3
echo "== fused"; CASES=0,0,1,2,3,4,5 timeout 120 python profiles/microbench/project_time.py
echo "== two launches"; B200_PROJECT_TWO_LAUNCHES=1 CASES=0,0,1,2,3,4,5 timeout 120 python profiles/microbench/project_time.py
