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
  tail -3 "gpurun_out/$log" | cut -c1-3000
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
step 300 r2k_bench8.log $TR --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5
step 300 r2k_bench8_staged.log env GCNB_BITTILE=0 $TR --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 5
step 200 r2k_dist_check8.log $TR --master-port 29515 scripts/dist_check_native.py
echo "== done"
