#!/bin/bash
# tiled projection route: parity + timing of both routes
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "project or projection or grid_sample" 2>&1 | tail -15
for r in two_pass tiled; do echo "== $r"; B200_PROJECT_ROUTE=$r CASES=${CASES:-5,0,1,2,3,4} python profiles/microbench/project_time.py; done 2>&1 | tee gpurun_out/r2_project_routes.txt
