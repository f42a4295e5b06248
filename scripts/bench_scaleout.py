"""Scale-out workload (BASELINE.json configs[4], north_star: ">= 5x 1-GPU throughput on 8 GPUs for a >= 1B-edge synthetic
graph"): a symmetric community graph with >= 1e9 CSR entries generated ROW-LOCALLY (every rank builds only its row
block, gcnb_synth_sym_rows), dense features, 2-layer GCN hidden 16, through the native row-partitioned engine.

  python scripts/bench_scaleout.py [--num-nodes 4000000 ...]                                  one GPU, whole graph
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 ... scripts/bench_scaleout.py [...]

Prints one JSON line (rank 0): ms/epoch (train_epoch + eval(2), CUDA events on the engine stream, max over ranks), mean
GraphSum time (slab exchange + SpMM), setup times.  torch.distributed only launches the ranks, ships the NCCL id and the
degree array, and reduces the timings."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-nodes", dest="n", type=int, default=4000000)
    ap.add_argument("--block", type=int, default=4000, help="community size")
    ap.add_argument("--intra", type=float, default=230.0)
    ap.add_argument("--inter", type=float, default=58.0)
    ap.add_argument("--reflect", type=int, default=2048)
    ap.add_argument("--sigma", type=float, default=1.0)
    ap.add_argument("--features", type=int, default=128)
    ap.add_argument("--classes", type=int, default=47)
    ap.add_argument("--hidden", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--seed", type=int, default=20240229)
    ap.add_argument("--inter-window", type=int, default=0, help="> 0: inter-community edges stay within this many rows (halo-exchange case)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    ge.load_package()
    gcnb = importlib.import_module("parallel_gcn_b200.binding")
    dmod = importlib.import_module("parallel_gcn_b200.dist")
    torch.cuda.set_device(local_rank)
    gcnb.device_check()
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = dmod.scaleout_record(rank, world, local_rank, dist, n=args.n, block=args.block, intra=args.intra, inter=args.inter,
                                reflect=args.reflect, sigma=args.sigma, features=args.features, classes=args.classes,
                                hidden=args.hidden, steps=args.steps, warmup=args.warmup, seed=args.seed, inter_window=args.inter_window)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
