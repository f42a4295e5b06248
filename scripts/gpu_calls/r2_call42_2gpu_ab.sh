#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for v in 0 1; do
  GCNB_BITS_FORK_LATE=$v timeout 200 $TR --master-port 2957$v bench.py --gpus 2 --steps 20 --warmup 5 --no-scaleout 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('late=$v', d['value'], r['mean_launch_us'], r['exchange_us'], r['product_us'], d['config']['final_train_loss'])
"
done
