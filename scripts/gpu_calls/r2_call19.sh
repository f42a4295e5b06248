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
  tail -${TAILN:-4} "gpurun_out/$log" | cut -c1-900
}
step 1500 r2s_gpu_tests.log python -m pytest tests -m gpu -q --durations=5
step 300 r2s_ncu_full.log ncu --set full --clock-control none --import-source on --kernel-name regex:"head_kernel|dense_feat_tn_mma|dense_feat_fwd_mma" --launch-skip 9 --launch-count 6 -o gpurun_out/r2s_head_dense -f python bench.py --no-cpu-baseline --no-extras --no-scaleout --steps 2 --warmup 1
step 200 r2s_ncu_cora.log ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s_cora_launches.csv python scripts/bench_configs.py cora
step 300 r2s_configs.log python scripts/bench_configs.py cora citeseer pubmed reddit600
echo "== done"
