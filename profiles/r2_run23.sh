#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "corr2d" 2>&1 | tail -3
