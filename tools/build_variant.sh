#!/bin/bash
# Build a named variant of libmodkit_b200.so for same-run A/B timing on the GPU box:
#   tools/build_variant.sh <name> [git-ref] [extra nvcc flags...]
# -> modkit_b200/_build/variants/<name>.so ; run with MODKIT_B200_LIB=<that path> python bench.py ...
set -e
NAME=$1; REF=${2:-WORK}; shift; shift || true
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/modkit_b200/_build/variants; mkdir -p $OUT
SRC=$ROOT
if [ "$REF" != "WORK" ]; then
  SRC=$(mktemp -d); git -C $ROOT archive $REF modkit_b200/csrc include | tar -x -C $SRC
fi
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=false \
  -Xcompiler -fPIC,-O3,-pthread,-ffp-contract=off "$@" -shared $SRC/modkit_b200/csrc/mkp_device.cu $SRC/modkit_b200/csrc/host/capi.cpp \
  -I$SRC/include -o $OUT/$NAME.so -lz -lcudart 2>&1 | grep -v "warning\|\^\|^$\|declared but never\|Remark" || true
ls -la $OUT/$NAME.so
