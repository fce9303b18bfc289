#!/bin/bash
bash profiles/ncu_kernel.sh "knn_grid_query_batched" 1 r2_knn python profiles/microbench/knn_time.py > /dev/null 2>&1
python profiles/ncu_summary.py gpurun_out/r2_knn_raw.csv; python profiles/ncu_stalls.py gpurun_out/r2_knn_raw.csv
python profiles/ncu_hot.py gpurun_out/r2_knn_source.csv knn_grid_query_batched
