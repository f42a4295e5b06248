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
  tail -${TAILN:-6} "gpurun_out/$log" | cut -c1-900
}
step 300 r2x_devbuild_tests.log python -m pytest tests/test_zz_bittile_gpu.py -q -x -k "device_buil or background_setup" --durations=5
step 300 r2x_bench.log env GCNB_SETUP_VERBOSE=1 python bench.py --no-cpu-baseline --no-extras --no-scaleout --steps 20 --warmup 5
step 900 r2x_gpu_tests.log python -m pytest tests -m gpu -q --durations=3
echo "== done"
