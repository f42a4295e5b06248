#!/usr/bin/env bash
# Round-2 first GPU call: everything that was written in round 1 without a GPU, in the order in which a failure is
# cheapest.  Run from the repo root on the GPU box, e.g.
#   gpurun --timeout 900 -- 'bash scripts/r2_checklist.sh > gpurun_out/r2_checklist.log 2>&1'
# Every step has its own timeout (a wedged mbarrier pipeline must not eat the budget) and its own log under gpurun_out/.
set -u
mkdir -p gpurun_out
step() {  # step <seconds> <log> <command...>
  local t=$1 log=$2
  shift 2
  echo "== $* (limit ${t}s) -> gpurun_out/$log"
  timeout -k 5 "$t" "$@" > "gpurun_out/$log" 2>&1
  echo "   rc=$?"
  tail -3 "gpurun_out/$log" | cut -c1-400
}

# 1. the validated first-generation bit-tile kernel + the default suite tail (sanity)
step 240 r2_bt_default.log python -m pytest tests/test_zz_bittile_gpu.py -m gpu -x -q
# 2. second-generation MMA kernel: unified stage barriers, 128-column tiles, 256-row items
step 60 r2_bt_small_unified.log env GCNB_BT_UNIFIED=1 python scripts/probe_bittile.py --stage small --out gpurun_out/r2_probe.jsonl
step 60 r2_bt_small_c128.log python scripts/probe_bittile.py --stage small --chunk 128 --out gpurun_out/r2_probe.jsonl
step 60 r2_bt_small_rb2.log python scripts/probe_bittile.py --stage small --rb 2 --out gpurun_out/r2_probe.jsonl
step 300 r2_bt_wide_tests.log env GCNB_TEST_BITTILE_WIDE=1 python -m pytest tests/test_zz_bittile_gpu.py -m gpu -x -q
# 3. timing on the bench graph, parts apart (pack | mma | remainder | add), each shape
step 120 r2_bt_g1_v0.log python scripts/probe_bittile.py --stage graph --scale 1 --iters 10 --staged 0 --out gpurun_out/r2_probe.jsonl
step 120 r2_bt_g1_unified.log env GCNB_BT_UNIFIED=1 python scripts/probe_bittile.py --stage graph --scale 1 --iters 10 --staged 0 --out gpurun_out/r2_probe.jsonl
step 120 r2_bt_g1_c128.log python scripts/probe_bittile.py --stage graph --scale 1 --iters 10 --staged 0 --chunk 128 --out gpurun_out/r2_probe.jsonl
step 120 r2_bt_g1_rb2.log python scripts/probe_bittile.py --stage graph --scale 1 --iters 10 --staged 0 --rb 2 --out gpurun_out/r2_probe.jsonl
step 120 r2_bt_g1_rb2_cap2.log env GCNB_BT_REM_CTAS=2 python scripts/probe_bittile.py --stage graph --scale 1 --iters 10 --staged 0 --rb 2 --out gpurun_out/r2_probe.jsonl
# 4. engine paths: bit tiles in the engine, background staging
step 300 r2_engine_optin.log env GCNB_TEST_BITTILE_ENGINE=1 GCNB_TEST_ASYNC_STAGE=1 python -m pytest tests/test_zz_bittile_gpu.py -m gpu -x -q -k "engine or background"
step 300 r2_ragged_engine.log env GCNB_TEST_RAGGED_ENGINE=1 python -m pytest tests/test_zz_bittile_gpu.py -m gpu -x -q -k ragged
# 4b. exact-split tcgen05 GEMM for the wide first layer (csrc/dense_tc.cu)
step 300 r2_dense_tc.log env GCNB_TEST_DENSE_TC=1 python -m pytest tests/test_zz_bittile_gpu.py -m gpu -x -q -k exact_split
step 300 r2_dense_tc_probe.log python scripts/probe_dense_tc.py --out gpurun_out/r2_probe_dense_tc.jsonl
step 600 r2_configs_default.log python scripts/bench_configs.py
step 600 r2_configs_dense_tc.log env GCNB_DENSE_TC=1 python scripts/bench_configs.py
step 600 r2_configs_dense_tc_bittile.log env GCNB_DENSE_TC=1 GCNB_BITTILE=1 python scripts/bench_configs.py
# 5. bench lines: default, background staging, bit tiles (+ the best shape from step 3 through GCNB_BT_CHUNK / GCNB_BT_RB)
step 300 r2_bench_default.log python bench.py --no-cpu-baseline
step 300 r2_bench_async.log env GCNB_ASYNC_STAGE=1 python bench.py --no-cpu-baseline
step 300 r2_bench_bittile.log env GCNB_BITTILE=1 python bench.py --no-cpu-baseline
step 300 r2_bench_bittile_rb2.log env GCNB_BITTILE=1 GCNB_BT_RB=2 python bench.py --no-cpu-baseline
# 5b. same-box A/B against the reference's OWN CUDA code (oracle/_ref/ref_gpu_bench, built by `make -C oracle`)
step 300 r2_ref_gpu_cora.log python scripts/bench_ref_gpu.py --dataset cora --epochs 100 --reps 5
step 300 r2_ref_gpu_citeseer.log python scripts/bench_ref_gpu.py --dataset citeseer --epochs 100 --reps 5
step 900 r2_ref_gpu_reddit.log python scripts/bench_ref_gpu.py --dataset reddit_shape --epochs 20 --reps 1 --timeout 600
# 5c. (needs `gpurun --gpus 2`) bit tiles in the row-partitioned engine: parity of 2 ranks against the oracle / single rank
#   GCNB_BITTILE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
#       scripts/dist_check_native.py
#   GCNB_BITTILE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
#       bench.py --gpus 2 --no-cpu-baseline
# 6. one full-set ncu capture of the MMA kernel of the best shape (edit the env / flags), after the runs above exited 0:
#   ncu --set full --clock-control none --import-source on -k regex:bt_mma -c 1 -o gpurun_out/r2_bt_mma \
#       python scripts/probe_bittile.py --stage graph --scale 1 --iters 1 --staged 0 --rb 2
echo "== done"
