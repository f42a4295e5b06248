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
step 1500 r2l_gpu_tests.log python -m pytest tests -m gpu -q --durations=5
step 400 r2l_configs.log python scripts/bench_configs.py cora citeseer pubmed
step 200 r2l_ref_gpu_cora.log python scripts/bench_ref_gpu.py --dataset cora --epochs 100 --reps 5
step 200 r2l_ref_gpu_citeseer.log python scripts/bench_ref_gpu.py --dataset citeseer --epochs 100 --reps 5
step 300 r2l_bench.log python bench.py --no-cpu-baseline --no-extras --no-scaleout --steps 20 --warmup 5
step 100 r2l_smoke.log python -c "import __graft_entry__ as g; g.smoke()"
echo "== done"
