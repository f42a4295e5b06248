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
  tail -4 "gpurun_out/$log" | cut -c1-1200
}
step 1500 r2m_gpu_tests.log python -m pytest tests -m gpu -q --durations=5
step 300 r2m_bench.log python bench.py --no-cpu-baseline --no-extras --no-scaleout --steps 20 --warmup 5
step 400 r2m_configs600.log python scripts/bench_configs.py reddit600
step 300 r2m_ncu_launches.log ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2m_bench_launches.csv python bench.py --no-cpu-baseline --no-extras --no-scaleout --steps 2 --warmup 1
echo "== done"
