#!/bin/bash
timeout 200 python profiles/microbench/voxel_time.py
bash profiles/ncu_kernel.sh "trilinear" 6 r2_voxel python profiles/microbench/voxel_time.py > /dev/null 2>&1
python profiles/ncu_summary.py gpurun_out/r2_voxel_raw.csv | cut -c1-200; python profiles/ncu_stalls.py gpurun_out/r2_voxel_raw.csv | grep -v "^    launch\|cycles"
