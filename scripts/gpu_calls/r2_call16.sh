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
  tail -6 "gpurun_out/$log" | cut -c1-900
}
step 300 r2p_new_tests.log python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_dropin_gpu.py -m gpu -q -k "output_head or tuning_sweep or reference_main" -x
step 300 r2p_sweep.log python scripts/bench_sweep.py --dataset cora --reps 1 --workers 1,2,4,8,16
step 300 r2p_bench.log python bench.py --no-cpu-baseline --no-extras --no-scaleout --steps 20 --warmup 5
GCNB_HEAD_CTAS=3 step 300 r2p_bench_ctas3.log python bench.py --no-cpu-baseline --no-extras --no-scaleout --steps 20 --warmup 5
echo "== done"
