#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
t0=$(date +%s)
timeout -k 5 600 $TR --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2ae_bench8.log 2>&1
echo "rc=$? ($(( $(date +%s) - t0 ))s)"
grep "^{" gpurun_out/r2ae_bench8.log | cut -c1-6000
grep -v "^{\|^\[W\|Warning\|^\*\*\*\|OMP_NUM" gpurun_out/r2ae_bench8.log | tail -15 | cut -c1-300
