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
  tail -${TAILN:-4} "gpurun_out/$log" | cut -c1-${CUT:-400}
}
step 900 r2ag_gpu_tests.log python -m pytest tests -m gpu -q --durations=3
step 900 r2ag_bench_full.log python bench.py --steps 20 --warmup 5
echo "== done"
