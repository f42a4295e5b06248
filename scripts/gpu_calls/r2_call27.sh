#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
step() {
  local t=$1 log=$2
  shift 2
  echo "== $* (limit ${t}s) -> gpurun_out/$log"
  local t0=$(date +%s)
  timeout -k 5 "$t" "$@" > "gpurun_out/$log" 2>&1
  echo "   rc=$? ($(( $(date +%s) - t0 ))s)"
  tail -${TAILN:-12} "gpurun_out/$log" | cut -c1-1200
}
step 300 r2z_rank_block.log python scripts/probe_rank_block.py --world 8 --rank 3
step 400 r2z_engine_tests.log python -m pytest tests/test_engine_gpu.py tests/test_configs_gpu.py tests/test_dropin_gpu.py -m gpu -q -x
echo "== done"
