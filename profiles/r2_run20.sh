#!/bin/bash
for r in two_pass tiled; do for s in "" 1; do echo "== $r serial=$s"; B200_LEVEL_SERIAL=$s B200_PROJECT_ROUTE=$r BATCH=74 python profiles/segment_times.py; done; done 2>&1 | grep -v Warn
