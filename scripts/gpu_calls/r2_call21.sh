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
  tail -${TAILN:-4} "gpurun_out/$log" | cut -c1-1000
}
step 1500 r2u_gpu_tests.log python -m pytest tests -m gpu -q --durations=3
step 300 r2u_bench.log python bench.py --no-cpu-baseline --no-extras --no-scaleout --steps 20 --warmup 5
step 300 r2u_ncu_full.log ncu --set full --clock-control none --import-source on --kernel-name regex:"head_tc_kernel" --launch-skip 3 --launch-count 2 -o gpurun_out/r2u_head_tc -f python bench.py --no-cpu-baseline --no-extras --no-scaleout --steps 2 --warmup 1
step 300 r2u_ncu_launches.log ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2u_bench_launches.csv python bench.py --no-cpu-baseline --no-extras --no-scaleout --steps 2 --warmup 1
echo "== done"
