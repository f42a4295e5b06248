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
  grep -v "^\[W\|Warning" "gpurun_out/$log" | tail -${TAILN:-12} | cut -c1-1200
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
TAILN=30 step 600 r2q_dist_check2.log $TR --master-port 29551 scripts/dist_check_native.py
TAILN=3 step 400 r2q_scaleout2_local.log $TR --master-port 29552 scripts/bench_scaleout.py --num-nodes 1000000 --inter-window 12000 --steps 5
GCNB_HALO=0 TAILN=3 step 400 r2q_scaleout2_local_nohalo.log $TR --master-port 29553 scripts/bench_scaleout.py --num-nodes 1000000 --inter-window 12000 --steps 5
echo "== done"
