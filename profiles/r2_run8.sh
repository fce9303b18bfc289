#!/bin/bash
for v in "" prej g8; do
  if [ -n "$v" ]; then export B200FLOW_LIB=$PWD/gpurun_variants/libb200flow_$v.so; fi
  echo "=== variant ${v:-default}"; timeout 120 python profiles/microbench/corr3d_time.py 2>&1 | tail -6
done
