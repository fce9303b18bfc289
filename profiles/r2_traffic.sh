#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"project_|corr2d_fwd|event_voxel|sample_point" -f -o gpurun_out/r2_traffic python profiles/r2_traffic.py > gpurun_out/r2_traffic.log 2>&1
ncu -i gpurun_out/r2_traffic.ncu-rep --page raw --csv > gpurun_out/r2_traffic_raw.csv
python profiles/ncu_summary.py gpurun_out/r2_traffic_raw.csv
