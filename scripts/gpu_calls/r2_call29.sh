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
  tail -${TAILN:-12} "gpurun_out/$log" | cut -c1-${CUT:-1200}
}
step 300 r2ab_rank_block.log python scripts/probe_rank_block.py --world 8 --rank 3
step 300 r2ab_rank_block4.log python scripts/probe_rank_block.py --world 4 --rank 1
CUT=300 TAILN=40 step 900 r2ab_bench_full.log env GCNB_SETUP_VERBOSE=1 python bench.py --steps 20 --warmup 5
echo "== done"
