#!/bin/bash
export B200_KNN_CTARGET_SMALLK=1
for c in 1 2 3 4 6 8; do
  echo "== small-k ctarget $c"; B200_KNN_CTARGET=$c timeout 300 python profiles/microbench/knn_time.py 2>&1 | grep -v Warn | sed -n 7,8p | tr '\n' ';'; echo
done
