#!/bin/bash
# usage: bash profiles/build_variant.sh <name> "<extra nvcc flags>"  -> gpurun_variants/libb200flow_<name>.so (kernel A/B experiments;
# select with B200FLOW_LIB=<path>)
set -e
NAME=$1; FLAGS=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $ROOT/gpurun_variants/build_$NAME
make -s -C $ROOT/rpeflow_b200/csrc -j8 OBJD=$ROOT/gpurun_variants/build_$NAME OUT=$ROOT/gpurun_variants/libb200flow_$NAME.so EXTRA="$FLAGS" > /dev/null
echo built gpurun_variants/libb200flow_$NAME.so
