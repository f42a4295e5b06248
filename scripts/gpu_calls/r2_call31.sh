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
  tail -${TAILN:-6} "gpurun_out/$log" | cut -c1-${CUT:-600}
}
step 900 r2ad_gpu_tests.log python -m pytest tests -m gpu -q --durations=3
step 200 r2ad_smoke.log python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')"
step 300 r2ad_bench_short.log python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-scaleout
step 400 r2ad_ncu_launches.log ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2ad_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-scaleout
step 300 r2ad_ref_arm.log python bench.py --impl reference --steps 1 --warmup 0
echo "== done"
