#!/bin/bash
# ncu --set full capture of selected kernels of the bench workload.
# Usage: tools/profile_one.sh <tag> <kernel-regex> [contig_len] [skip] [count]
TAG=$1; K=$2; LEN=${3:-8000000}; SKIP=${4:-4}; CNT=${5:-1}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$K" -s $SKIP -c $CNT -o gpurun_out/${TAG} -f python bench.py --contig-len $LEN --steps 2 --warmup 3 --skip-cpu > gpurun_out/${TAG}.log 2>&1
ls -la gpurun_out/${TAG}.ncu-rep
