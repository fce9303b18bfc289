# usage: bash profiles/ncu_kernel.sh <kernel-name regex> <launch count> <output stem> <command...>
# one `ncu --set full` capture of the kernels matching the regex (run under gpurun; writes gpurun_out/<stem>*):
# raw metrics, the per-source-line page (stall sampling) and the summary table.
set -e
REGEX=$1; COUNT=$2; STEM=$3; shift 3
ncu --set full --clock-control none --import-source on -k regex:$REGEX -c $COUNT -f -o gpurun_out/$STEM "$@" > gpurun_out/$STEM.log 2>&1
ncu -i gpurun_out/$STEM.ncu-rep --page raw --csv > gpurun_out/${STEM}_raw.csv
ncu -i gpurun_out/$STEM.ncu-rep --page source --csv > gpurun_out/${STEM}_source.csv 2>/dev/null || true
python profiles/ncu_summary.py gpurun_out/${STEM}_raw.csv
python profiles/ncu_stalls.py gpurun_out/${STEM}_raw.csv
