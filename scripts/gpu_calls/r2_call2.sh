#!/usr/bin/env bash
# Round-2 GPU call 2: full GPU suite with every former opt-in test un-gated, ELL remainder probe, default bench line.
set -u
mkdir -p gpurun_out
step() {
  local t=$1 log=$2
  shift 2
  echo "== $* (limit ${t}s) -> gpurun_out/$log"
  local t0=$(date +%s)
  timeout -k 5 "$t" "$@" > "gpurun_out/$log" 2>&1
  echo "   rc=$? ($(( $(date +%s) - t0 ))s)"
  tail -4 "gpurun_out/$log" | cut -c1-700
}
P=gpurun_out/r2b_probe.jsonl
step 1500 r2b_gpu_tests.log python -m pytest tests -m gpu -q --durations=15
step 100 r2b_bt_ell.log python scripts/probe_bittile.py --stage graph --scale 1 --iters 10 --staged 0 --out $P
step 100 r2b_bt_noell.log env GCNB_BT_ELL=0 python scripts/probe_bittile.py --stage graph --scale 1 --iters 10 --staged 0 --out $P
step 100 r2b_bt_ell_c1.log env GCNB_BT_REM_CTAS=1 python scripts/probe_bittile.py --stage graph --scale 1 --iters 10 --staged 0 --out $P
step 100 r2b_bt_ell_c3.log env GCNB_BT_REM_CTAS=3 python scripts/probe_bittile.py --stage graph --scale 1 --iters 10 --staged 0 --out $P
step 100 r2b_bt_ell_rb2.log python scripts/probe_bittile.py --stage graph --scale 1 --iters 10 --staged 0 --rb 2 --out $P
step 400 r2b_bench.log python bench.py
step 300 r2b_bench_ref.log python bench.py --impl reference --steps 2 --warmup 1
step 300 r2b_ncu_ell.log ncu --set full --clock-control none --import-source on -k regex:ell_gather -c 1 -f -o gpurun_out/r2b_ell python scripts/probe_bittile.py --stage graph --scale 1 --iters 1 --staged 0 --out gpurun_out/r2b_probe_ncu.jsonl
step 300 r2b_ncu_c128.log ncu --set full --clock-control none --import-source on -k regex:bt_mma -c 1 -f -o gpurun_out/r2b_bt_mma_c128 python scripts/probe_bittile.py --stage graph --scale 1 --iters 1 --staged 0 --out gpurun_out/r2b_probe_ncu.jsonl
echo "== done"
