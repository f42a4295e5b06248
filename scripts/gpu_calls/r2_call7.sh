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
  tail -4 "gpurun_out/$log" | cut -c1-1500
}
step 1500 r2g_gpu_tests.log env GCNB_TEST_DENSE_TC=1 python -m pytest tests -m gpu -q --durations=8
step 200 r2g_dense_tc_probe.log python scripts/probe_dense_tc.py --out gpurun_out/r2g_probe_dense_tc.jsonl
step 400 r2g_configs.log python scripts/bench_configs.py reddit600
step 400 r2g_configs_tc.log env GCNB_DENSE_TC=1 python scripts/bench_configs.py reddit600
step 900 r2g_bench_full.log python bench.py --steps 20 --warmup 5
echo "== done"
