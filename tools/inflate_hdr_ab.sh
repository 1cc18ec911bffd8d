#!/bin/bash
# Same-box A/B of k_inflate's header batching (lanes of a warp that run a deflate block header together; 1 = the old behaviour).
for b in 1 4 8 12 16 24; do
  export MKP_INFLATE_HDR_BATCH=$b
  echo "== hdr_batch $b"; timeout 300 python tools/ingest_bench.py 64444167 2 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(d[k],1) for k in ('h2d_ms','inflate_ms','walk_ms')}, round(d['open_wall_s'],3), round(d['inflate_in_GBps'],2))"
done
