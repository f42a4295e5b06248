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
  tail -4 "gpurun_out/$log" | cut -c1-2500
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
step 400 r2i_test2gpu.log python -m pytest tests/test_configs_gpu.py -m gpu -q -k two_gpus
step 400 r2i_bench2.log env GCNB_SETUP_VERBOSE=1 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5
step 300 r2i_bench2_staged.log env GCNB_BITTILE=0 $TR --master-port 29514 bench.py --gpus 2 --no-scaleout --steps 20 --warmup 5
echo "== done"
