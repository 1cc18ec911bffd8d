#!/bin/bash
# Wall-clock of the drop-in CLI (BAM file -> bedMethyl file) on the bench workload with both ingest paths (GPU inflate +
# slicing vs host zlib), next to the CPU oracle on a window.
set -e
D=${TMPDIR:-/dev/shm}/mkb200_cli_$$; mkdir -p $D; trap "rm -rf $D" EXIT
LEN=${1:-64444167}
tools/_build/synth_modbam --out $D/w --contig syn1:$LEN --coverage 50 --mods hm --level 1 --threads 64 > $D/gen.json
ls -la $D/w.bam | awk '{print "bam_bytes", $5}'
for mode in device host device; do
  EXTRA=""; [ $mode = host ] && EXTRA="--host-ingest"
  T0=$(date +%s.%N); modkit_b200/_build/modkit pileup -t 64 --cpg --ref $D/w.fa $D/w.bam $D/gpu_$mode.bed --stats-json gpurun_out/cli_fullsize_stats_$mode.json $EXTRA; T1=$(date +%s.%N); echo "cli_wall_s[$mode] $(python -c "print($T1 - $T0)")"
  cat gpurun_out/cli_fullsize_stats_$mode.json
done
cmp $D/gpu_device.bed $D/gpu_host.bed && echo "device == host ingest ($(wc -l < $D/gpu_device.bed) rows)"
T0=$(date +%s.%N); oracle/_build/modkit_oracle pileup -t 128 --cpg --ref $D/w.fa --region syn1:0-8000000 $D/w.bam $D/cpu.bed; T1=$(date +%s.%N); echo "oracle_window_wall_s $(python -c "print($T1 - $T0)")"
modkit_b200/_build/modkit pileup -t 64 --quiet --cpg --ref $D/w.fa --region syn1:0-8000000 $D/w.bam $D/gpu_win.bed
cmp $D/cpu.bed $D/gpu_win.bed && echo "window parity OK ($(wc -l < $D/cpu.bed) rows)"
