#!/bin/bash
for v in "" knn_mb5 knn_mb6; do
  lib=gpurun_variants/libb200flow_$v.so; [ -z "$v" ] && lib=rpeflow_b200/libb200flow.so
  echo "== ${v:-base}"; B200FLOW_LIB=$lib timeout 300 python profiles/microbench/knn_time.py 2>&1 | grep -v Warn | head -6
done
