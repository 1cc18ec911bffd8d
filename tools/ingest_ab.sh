#!/bin/bash
# Same-box timing of library variants on the device ingest: tools/ingest_ab.sh <contig_len> <variant|cur> ...
L=$1; shift
for v in "$@"; do
  if [ $v = cur ]; then unset MODKIT_B200_LIB; else export MODKIT_B200_LIB=$PWD/modkit_b200/_build/variants/$v.so; fi
  echo "== $v"; timeout 600 python tools/ingest_bench.py $L 2 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(d[k],1) for k in ('h2d_ms','inflate_ms','walk_ms')}, round(d['open_wall_s'],3))"
done
