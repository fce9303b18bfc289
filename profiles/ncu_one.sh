# usage: bash profiles/ncu_one.sh <PROFILE_ONLY tag> <output stem>   (run under gpurun; writes gpurun_out/<stem>*)
set -e
PROFILE_ONLY=$1 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/$2 python profiles/profile_ops.py > gpurun_out/$2.log 2>&1
ncu -i gpurun_out/$2.ncu-rep --page raw --csv > gpurun_out/$2_raw.csv
ncu -i gpurun_out/$2.ncu-rep --page source --csv > gpurun_out/$2_source.csv 2>/dev/null || true
python profiles/ncu_summary.py gpurun_out/$2_raw.csv
