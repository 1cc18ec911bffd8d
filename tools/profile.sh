#!/bin/bash
# ncu evidence for the bench workload (B200_PROFILING.md recipe). Usage: tools/profile.sh <tag> [contig_len]
# Writes gpurun_out/<tag>_launches.csv (every launch, device time) and gpurun_out/<tag>_{decode,bases}.ncu-rep (--set full).
TAG=${1:-prof}
LEN=${2:-8000000}
mkdir -p gpurun_out
CMD="python bench.py --contig-len $LEN --steps 2 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_resolve -s 4 -c 1 -o gpurun_out/${TAG}_resolve -f $CMD > gpurun_out/${TAG}_resolve.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_parse -s 4 -c 1 -o gpurun_out/${TAG}_parse -f $CMD > gpurun_out/${TAG}_parse.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_count_bases -s 3 -c 1 -o gpurun_out/${TAG}_bases -f $CMD > gpurun_out/${TAG}_bases.log 2>&1
ls -la gpurun_out/
