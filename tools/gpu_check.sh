#!/bin/bash
# Runs on the GPU box (gpurun): the GPU test suite with per-test time limits and a live log under gpurun_out/.
# usage: tools/gpu_check.sh [pytest args...]
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=180 --timeout-method=thread "$@" 2>&1 | tee gpurun_out/tests.log | tail -25
