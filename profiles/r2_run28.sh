#!/bin/bash
# PointConv v2 timing experiments (variants built with -DP2_EXP_*: results wrong, timing only)
for v in pc_NOCOPYNOMMANOFENCE pc_NOCOPYNOMMANOGATHER pc_NOGATHER pc_NOFENCE; do
  lib=gpurun_variants/libb200flow_$v.so; [ -z "$v" ] && lib=rpeflow_b200/libb200flow.so
  echo "== ${v:-base}"; B200FLOW_LIB=$lib ONLY=down0,est1_1024,est2_1024 timeout 200 python profiles/microbench/pointconv_time.py 2>&1 | grep -v Warn | head -3
done
