#!/bin/bash
timeout 200 python profiles/microbench/voxel_time.py
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -k "voxel or event or trilinear" 2>&1 | tail -2
bash profiles/r2_measure1.sh
