#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "voxel or trilinear or event" 2>&1 | tail -3
timeout 300 python profiles/microbench/voxel_time.py 2>&1 | grep -v Warn
