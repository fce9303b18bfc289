#!/bin/bash
mkdir -p gpurun_out
B200_PROJECT_ROUTE=tiled CASES=2 bash profiles/ncu_kernel.sh "project_" 7 r2_project_tile python profiles/microbench/project_time.py
