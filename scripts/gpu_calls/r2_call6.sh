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
  tail -4 "gpurun_out/$log" | cut -c1-1200
}
step 1500 r2f_gpu_tests.log python -m pytest tests -m gpu -q --durations=10
step 300 r2f_bench_rb2.log env GCNB_BT_RB=2 python bench.py --no-cpu-baseline --no-extras --no-scaleout --steps 20 --warmup 5
step 300 r2f_bench_c128.log python bench.py --no-cpu-baseline --no-extras --no-scaleout --steps 20 --warmup 5
step 900 r2f_bench_full.log python bench.py --steps 20 --warmup 5
echo "== done"
