#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "project or projection or grid_sample" 2>&1 | tail -3
B200_PROJECT_ROUTE=tiled CASES=2,2,0,1,3,4 python profiles/microbench/project_time.py 2>&1
BATCH=74 python profiles/segment_times.py 2>&1 | grep -v Warn
