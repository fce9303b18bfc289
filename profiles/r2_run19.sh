#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "project or projection or grid_sample" 2>&1 | tail -5
for r in tiled; do echo "== $r"; B200_PROJECT_ROUTE=$r CASES=2,2,0,1,3,4 python profiles/microbench/project_time.py; done 2>&1
for r in two_pass tiled; do echo "== $r"; B200_PROJECT_ROUTE=$r BATCH=74 python profiles/segment_times.py; done 2>&1 | grep -v Warn
