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
  tail -5 "gpurun_out/$log" | cut -c1-1500
}
step 1500 r2j_gpu_tests.log python -m pytest tests -m gpu -q --durations=5
step 400 r2j_configs.log python scripts/bench_configs.py
step 200 r2j_ncu_cora.log ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2j_cora_launches.csv python scripts/bench_configs.py cora
step 300 r2j_ncu_graphsum.log ncu --set full --clock-control none --import-source on -k regex:bt_pack\|bt_mma\|ell_gather -c 3 -f -o gpurun_out/r2j_graphsum python scripts/probe_bittile.py --stage graph --scale 1 --iters 1 --staged 0 --out gpurun_out/r2j_probe_ncu.jsonl
echo "== done"
