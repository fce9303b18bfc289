set -e
PROFILE_ONLY=corr2d ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/corr2d_v2 python profiles/profile_ops.py > gpurun_out/corr2d_v2.log 2>&1
ncu -i gpurun_out/corr2d_v2.ncu-rep --page raw --csv > gpurun_out/corr2d_v2_raw.csv
ncu -i gpurun_out/corr2d_v2.ncu-rep --page source --csv > gpurun_out/corr2d_v2_source.csv 2>/dev/null || true
python profiles/ncu_summary.py gpurun_out/corr2d_v2_raw.csv
