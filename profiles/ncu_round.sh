# Round artefacts (run under gpurun): (1) launch list of one bench step, (2) --set full capture of every hot-path kernel.
set -e
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$1.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --eager > gpurun_out/bench_under_ncu_$1.log 2>&1 || true
ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/ops_$1 \
    python profiles/profile_ops.py > gpurun_out/ops_$1.log 2>&1
ncu -i gpurun_out/ops_$1.ncu-rep --page raw --csv > gpurun_out/ops_$1_raw.csv
python profiles/ncu_summary.py gpurun_out/ops_$1_raw.csv > gpurun_out/ops_$1_summary.md
cat gpurun_out/ops_$1_summary.md
