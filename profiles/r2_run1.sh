#!/bin/bash
# round 2, GPU call 1: full GPU test-suite (incl. the model drop-in test) + one default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,memory.total --format=csv > gpurun_out/r2_smi1.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_pytest1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
echo "bench rc=$?" >> gpurun_out/r2_bench1.err
tail -5 gpurun_out/r2_pytest1.log
