#!/bin/bash
for v in 20 30 45 1000; do
  lib=gpurun_variants/libb200flow_tight$v.so; [ $v = 20 ] && lib=rpeflow_b200/libb200flow.so
  echo "== KG_TIGHT_X10=$v"; B200FLOW_LIB=$lib python profiles/microbench/knn_time.py 2>&1 | grep -v Warn | head -4
done
