#!/bin/bash
# Round evidence for profiles/: (1) launch list of a short bench run (per-launch gpu__time_duration), (2) one `--set full`
# capture of every per-read kernel + the ingest kernels, all at a reduced contig (default 8 Mb) so ncu's replays stay short.
# Usage: tools/profile.sh <tag> [contig_len]
TAG=${1:-rXX}; LEN=${2:-8000000}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --contig-len $LEN --steps 2 --warmup 3 --skip-cpu > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_parse|k_resolve|k_count_bases|k_count_calls|k_rows" -s 14 -c 7 \
    -o gpurun_out/${TAG}_pileup -f python bench.py --contig-len $LEN --steps 2 --warmup 3 --skip-cpu > gpurun_out/${TAG}_pileup.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_inflate|k_walk|k_slice" -c 7 \
    -o gpurun_out/${TAG}_ingest -f python tools/ingest_bench.py $LEN 1 > gpurun_out/${TAG}_ingest.log 2>&1
ls -la gpurun_out/${TAG}_*
