#!/usr/bin/env bash
# 2-GPU call: native partitioned engine, window-staged (default) vs bit tiles (GCNB_BITTILE=1): parity + bench lines
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
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
step 400 r2e_dist_check.log $TR --master-port 29511 scripts/dist_check_native.py
step 400 r2e_dist_check_bt.log env GCNB_BITTILE=1 $TR --master-port 29512 scripts/dist_check_native.py
step 300 r2e_bench2.log env GCNB_SETUP_VERBOSE=1 $TR --master-port 29513 bench.py --gpus 2 --no-scaleout
step 300 r2e_bench2_bt.log env GCNB_SETUP_VERBOSE=1 GCNB_BITTILE=1 $TR --master-port 29514 bench.py --gpus 2 --no-scaleout
step 300 r2e_test2gpu.log python -m pytest tests/test_configs_gpu.py -m gpu -q -k two_gpus
echo "== done"
