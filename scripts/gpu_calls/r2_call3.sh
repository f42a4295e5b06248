#!/usr/bin/env bash
# Round-2 GPU call 3: merge-by-reduction, drop-in execution tests, launch list of the bench step.
set -u
mkdir -p gpurun_out
step() {
  local t=$1 log=$2
  shift 2
  echo "== $* (limit ${t}s) -> gpurun_out/$log"
  local t0=$(date +%s)
  timeout -k 5 "$t" "$@" > "gpurun_out/$log" 2>&1
  echo "   rc=$? ($(( $(date +%s) - t0 ))s)"
  tail -4 "gpurun_out/$log" | cut -c1-900
}
P=gpurun_out/r2c_probe.jsonl
step 1500 r2c_gpu_tests.log python -m pytest tests -m gpu -q --durations=10
step 100 r2c_bt_merge.log python scripts/probe_bittile.py --stage graph --scale 1 --iters 10 --staged 0 --out $P
step 100 r2c_bt_nomerge.log env GCNB_BT_MERGE=0 python scripts/probe_bittile.py --stage graph --scale 1 --iters 10 --staged 0 --out $P
step 300 r2c_bench.log python bench.py --no-cpu-baseline --no-extras --steps 20 --warmup 5
step 400 r2c_configs.log python scripts/bench_configs.py
step 300 r2c_ncu_launches.log ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c_bench_launches.csv python bench.py --no-cpu-baseline --no-extras --steps 2 --warmup 1
echo "== done"
