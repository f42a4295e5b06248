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
  grep -v "^\[W\|Warning\|^\*\*\*\|OMP_NUM" "gpurun_out/$log" | tail -${TAILN:-12} | cut -c1-${CUT:-1200}
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
TAILN=6 step 600 r2aa_dist_check2.log $TR --master-port 29551 scripts/dist_check_native.py
TAILN=12 CUT=4000 step 400 r2aa_bench2.log env GCNB_SETUP_VERBOSE=1 $TR --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 5 --no-scaleout
echo "== done"
