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
  grep -v "^{" "gpurun_out/$log" | tail -${TAILN:-30} | cut -c1-300
}
step 300 r2y_bench.log env GCNB_SETUP_VERBOSE=1 python bench.py --no-cpu-baseline --no-extras --no-scaleout --steps 20 --warmup 5
step 300 r2y_bench_hostbuild.log env GCNB_SETUP_VERBOSE=1 GCNB_BT_DEVICE_BUILD=0 python bench.py --no-cpu-baseline --no-extras --no-scaleout --steps 20 --warmup 5
echo "== done"
