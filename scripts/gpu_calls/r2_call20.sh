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
  tail -${TAILN:-4} "gpurun_out/$log" | cut -c1-1200
}
TAILN=12 step 300 r2t_head_test.log python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "output_head"
step 400 r2t_cfg_tests.log python -m pytest tests/test_configs_gpu.py tests/test_engine_gpu.py -m gpu -q -x
GCNB_SETUP_VERBOSE=1 step 300 r2t_bench.log python bench.py --no-cpu-baseline --no-extras --no-scaleout --steps 20 --warmup 5
grep "\[setup\]\|\[bittile\]\|bt\]" gpurun_out/r2t_bench.log | head -40
GCNB_HEAD_TC=0 step 300 r2t_bench_fma.log python bench.py --no-cpu-baseline --no-extras --no-scaleout --steps 20 --warmup 5
step 300 r2t_configs600.log python scripts/bench_configs.py reddit600
echo "== done"
