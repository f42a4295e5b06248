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
  tail -3 "gpurun_out/$log" | cut -c1-900
}
P=gpurun_out/r2d_probe.jsonl
step 100 r2d_bt_c128.log python scripts/probe_bittile.py --stage graph --scale 1 --iters 10 --staged 0 --out $P
step 100 r2d_bt_rb2.log python scripts/probe_bittile.py --stage graph --scale 1 --iters 10 --staged 0 --rb 2 --out $P
step 100 r2d_bt_c64.log python scripts/probe_bittile.py --stage graph --scale 1 --iters 10 --staged 0 --chunk 64 --out $P
step 300 r2d_bench.log python bench.py --no-cpu-baseline --no-extras --no-scaleout --steps 20 --warmup 5
step 300 r2d_ncu_launches.log ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2d_bench_launches.csv python bench.py --no-cpu-baseline --no-extras --no-scaleout --steps 2 --warmup 1
step 600 r2d_scaleout1.log python scripts/bench_scaleout.py
echo "== done"
