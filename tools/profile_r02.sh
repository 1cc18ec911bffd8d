#!/bin/bash
# Round-2 evidence for profiles/ (run on the GPU box through gpurun; everything lands in gpurun_out/):
#  1. launch list of the bench command (pass kernels): ncu --metrics gpu__time_duration.sum --clock-control none
#  2. DRAM traffic of one whole pass over the genome (the launches of the first warm-up pass): dram__bytes_read/write.sum
#  3. `--set full --import-source on` of every pass kernel, the fused kernel and the ingest kernels on a 16 Mb chunk
# Usage: tools/profile_r02.sh [tag]
TAG=${1:-r02}
K="regex:k_parse|k_resolve|k_count_calls|k_count_bases|k_rows"
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/${TAG}_launches.log 2>&1
# 8 pieces: the sampling shortcut launches 3 matching kernels per piece (24), then every pass launches 7 per piece (56)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "$K" -s 24 -c 56 --csv \
    --log-file gpurun_out/${TAG}_traffic.csv python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/${TAG}_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 14 -c 7 -o gpurun_out/${TAG}_pileup -f \
    python tools/kbench.py --len 16000000 --steps 2 cur > gpurun_out/${TAG}_pileup.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_pileup_fused" -s 2 -c 1 -o gpurun_out/${TAG}_fused -f \
    python tools/kbench.py --len 16000000 --steps 2 fused > gpurun_out/${TAG}_fused.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_inflate|k_walk|k_slice" -c 7 -o gpurun_out/${TAG}_ingest -f \
    python tools/ingest_bench.py 16000000 1 > gpurun_out/${TAG}_ingest.log 2>&1
ls -la gpurun_out/${TAG}_*
