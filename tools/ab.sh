#!/bin/bash
# Same-box A/B timing of library variants (tools/build_variant.sh): tools/ab.sh [--tests] <variant|cur> ...
# "cur" = modkit_b200/_build/libmodkit_b200.so. The workload is generated once and reused.
if [ "$1" = "--tests" ]; then shift; python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
mkdir -p gpurun_out
for v in "$@"; do
  if [ $v = cur ]; then unset MODKIT_B200_LIB; else export MODKIT_B200_LIB=$PWD/modkit_b200/_build/variants/$v.so; fi
  python bench.py --steps 10 --warmup 3 --skip-cpu --workdir /dev/shm/mkb_ab > gpurun_out/ab_$v.json 2>gpurun_out/ab_$v.err || tail -5 gpurun_out/ab_$v.err
  python - <<PY
import json; d=json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1]); print("$v", round(d["ms_per_step"],3), {k: round(x,3) for k,x in d.get("stage_ms",{}).items()}, "e2e %.4g" % d["e2e"]["value"])
PY
done
