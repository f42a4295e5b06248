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
  tail -${TAILN:-4} "gpurun_out/$log" | cut -c1-700
}
step 1500 r2w_gpu_tests.log python -m pytest tests -m gpu -q --durations=3
step 300 r2w_configs.log python scripts/bench_configs.py cora citeseer pubmed
step 300 r2w_bench.log python bench.py --no-cpu-baseline --no-extras --no-scaleout --steps 20 --warmup 5
step 300 r2w_sweep.log python scripts/bench_sweep.py --dataset cora --reps 1 --workers 1,8
step 200 r2w_ncu_cora.log ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w_cora_launches.csv python scripts/bench_configs.py cora
echo "== done"
