timeout 300 python -m pytest tests/test_gpu_ingest.py -x -q 2>&1 | tail -2
for v in s1 cur s8; do
  if [ $v = cur ]; then unset MODKIT_B200_LIB; else export MODKIT_B200_LIB=$PWD/modkit_b200/_build/variants/$v.so; fi
  echo "== $v"; timeout 600 python tools/ingest_bench.py 32000000 2 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(d[k],1) for k in ('h2d_ms','inflate_ms','walk_ms')}, round(d['open_wall_s'],3))"
done
unset MODKIT_B200_LIB
ncu --set full --clock-control none --import-source on -k regex:k_inflate -c 1 -o gpurun_out/r1r_inflate -f python tools/ingest_bench.py 16000000 1 > gpurun_out/r1r_inflate.log 2>&1
