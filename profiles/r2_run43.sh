#!/bin/bash
for c in 4 6 8 10 12 16; do
  echo "== ctarget $c"; B200_KNN_CTARGET=$c timeout 300 python profiles/microbench/knn_time.py 2>&1 | grep -v Warn | head -6 | tr '\n' ';'; echo
done
