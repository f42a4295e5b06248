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
  grep -v "^\[W\|Warning\|^\*\*\*\|NCCL version" "gpurun_out/$log" | tail -${TAILN:-3} | cut -c1-2500
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
step 500 r2r_bench8.log $TR --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5
step 300 r2r_scaleout8_local.log $TR --master-port 29562 scripts/bench_scaleout.py --inter-window 12000 --steps 5
GCNB_HALO=0 step 300 r2r_scaleout8_local_nohalo.log $TR --master-port 29563 scripts/bench_scaleout.py --inter-window 12000 --steps 5
echo "== done"
