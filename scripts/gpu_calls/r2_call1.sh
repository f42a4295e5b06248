#!/usr/bin/env bash
# Round-2 GPU call 1: the r1 checklist, trimmed to fit ~25 GPU-minutes, + ncu captures of the MMA kernels.
set -u
mkdir -p gpurun_out
step() {
  local t=$1 log=$2
  shift 2
  echo "== $* (limit ${t}s) -> gpurun_out/$log"
  local t0=$(date +%s)
  timeout -k 5 "$t" "$@" > "gpurun_out/$log" 2>&1
  echo "   rc=$? ($(( $(date +%s) - t0 ))s)"
  tail -3 "gpurun_out/$log" | cut -c1-600
}
P=gpurun_out/r2_probe.jsonl
step 200 r2_bt_default.log python -m pytest tests/test_zz_bittile_gpu.py -m gpu -x -q
step 60 r2_bt_small_unified.log env GCNB_BT_UNIFIED=1 python scripts/probe_bittile.py --stage small --out $P
step 60 r2_bt_small_c128.log python scripts/probe_bittile.py --stage small --chunk 128 --out $P
step 60 r2_bt_small_rb2.log python scripts/probe_bittile.py --stage small --rb 2 --out $P
step 240 r2_bt_wide_tests.log env GCNB_TEST_BITTILE_WIDE=1 python -m pytest tests/test_zz_bittile_gpu.py -m gpu -x -q
step 100 r2_bt_g1_v0.log python scripts/probe_bittile.py --stage graph --scale 1 --iters 10 --staged 0 --out $P
step 100 r2_bt_g1_unified.log env GCNB_BT_UNIFIED=1 python scripts/probe_bittile.py --stage graph --scale 1 --iters 10 --staged 0 --out $P
step 100 r2_bt_g1_c128.log python scripts/probe_bittile.py --stage graph --scale 1 --iters 10 --staged 0 --chunk 128 --out $P
step 100 r2_bt_g1_rb2.log python scripts/probe_bittile.py --stage graph --scale 1 --iters 10 --staged 0 --rb 2 --out $P
step 240 r2_engine_optin.log env GCNB_TEST_BITTILE_ENGINE=1 GCNB_TEST_ASYNC_STAGE=1 python -m pytest tests/test_zz_bittile_gpu.py -m gpu -x -q -k "engine or background"
step 240 r2_ragged_engine.log env GCNB_TEST_RAGGED_ENGINE=1 python -m pytest tests/test_zz_bittile_gpu.py -m gpu -x -q -k ragged
step 240 r2_dense_tc.log env GCNB_TEST_DENSE_TC=1 python -m pytest tests/test_zz_bittile_gpu.py -m gpu -x -q -k exact_split
step 200 r2_dense_tc_probe.log python scripts/probe_dense_tc.py --out gpurun_out/r2_probe_dense_tc.jsonl
step 400 r2_configs_default.log python scripts/bench_configs.py
step 400 r2_configs_dense_tc.log env GCNB_DENSE_TC=1 python scripts/bench_configs.py
step 200 r2_bench_default.log python bench.py --no-cpu-baseline
step 200 r2_bench_async.log env GCNB_ASYNC_STAGE=1 python bench.py --no-cpu-baseline
step 200 r2_bench_bittile.log env GCNB_BITTILE=1 python bench.py --no-cpu-baseline
step 200 r2_ref_gpu_cora.log python scripts/bench_ref_gpu.py --dataset cora --epochs 100 --reps 5
step 200 r2_ref_gpu_citeseer.log python scripts/bench_ref_gpu.py --dataset citeseer --epochs 100 --reps 5
step 600 r2_ref_gpu_reddit.log python scripts/bench_ref_gpu.py --dataset reddit_shape --epochs 20 --reps 1 --timeout 400
# ncu full-set captures of the MMA kernels (first generation and 256-row items)
step 300 r2_ncu_bt_v0.log ncu --set full --clock-control none --import-source on -k regex:bt_mma -c 1 -f -o gpurun_out/r2_bt_mma_v0 python scripts/probe_bittile.py --stage graph --scale 1 --iters 1 --staged 0 --out gpurun_out/r2_probe_ncu.jsonl
step 300 r2_ncu_bt_rb2.log ncu --set full --clock-control none --import-source on -k regex:bt_mma -c 1 -f -o gpurun_out/r2_bt_mma_rb2 python scripts/probe_bittile.py --stage graph --scale 1 --iters 1 --staged 0 --rb 2 --out gpurun_out/r2_probe_ncu.jsonl
echo "== done"
