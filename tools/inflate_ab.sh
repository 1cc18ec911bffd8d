#!/bin/bash
# Same-box A/B of the BGZF inflate kernel: table placement / resident warps (env) and register budgets (library variants).
for cfg in "cur WARPS=22" "I28 WARPS=28" "I32 WARPS=32" "I32 WARPS=24"; do
  set -- $cfg
  unset MKP_INFLATE_SMEM MKP_INFLATE_WARPS MODKIT_B200_LIB
  [ $1 != cur ] && export MODKIT_B200_LIB=$PWD/modkit_b200/_build/variants/$1.so
  export MKP_INFLATE_$2
  echo "== $cfg"; timeout 300 python tools/ingest_bench.py 64444167 2 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(d[k],1) for k in ('h2d_ms','inflate_ms','walk_ms')}, round(d['open_wall_s'],3), round(d['inflate_in_GBps'],2))"
done
