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
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    ge.load_package()
    gcnb = importlib.import_module("parallel_gcn_b200.binding")
    eng = importlib.import_module("parallel_gcn_b200.engine")
    dmod = importlib.import_module("parallel_gcn_b200.dist")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    gcnb.device_check()
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    n = args.n
    B = dmod.block_rows(n, world)
    r0, r1 = min(n, rank * B), min(n, (rank + 1) * B)
    rows = r1 - r0
    t0 = time.perf_counter()
    g_indptr, g_indices = eng.synth_sym_rows(n, r0, rows, args.block, args.intra, args.inter, args.reflect, args.sigma, args.seed)
    t_graph = time.perf_counter() - t0
    deg_local = np.diff(g_indptr.astype(np.int64)).astype(np.uint32)
    if world > 1:
        pad = torch.zeros(B, dtype=torch.int32, device=dev)
        pad[:rows] = torch.from_numpy(deg_local.view(np.int32)).to(dev)
        allpad = torch.empty(world * B, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(allpad, pad)
        deg_global = allpad[:n].cpu().numpy().view(np.uint32)
        nnz_t = torch.tensor([len(g_indices)], dtype=torch.int64, device=dev)
        dist.all_reduce(nnz_t)
        nnz_global = int(nnz_t.item())
    else:
        deg_global, nnz_global = deg_local, len(g_indices)
    gv = eng.synth_graph_values(g_indptr, g_indices, r0, deg_global)
    f_indptr, f_indices, f_value = eng.synth_dense_features_uniform(rows, args.features, args.seed, r0 * args.features)
    label_all, split_all = eng.synth_labels(n, args.classes, seed=args.seed)
    t_gen = time.perf_counter() - t0
    part = dict(n_global=n, block=B, r0=r0, r1=r1, n_local=rows, g_indptr=g_indptr, g_indices=g_indices, graph_value=gv,
                f_indptr=f_indptr, f_indices=f_indices, f_value=f_value, f_elem_offset=r0 * args.features,
                label=np.ascontiguousarray(label_all[r0:r1]), split=np.ascontiguousarray(split_all[r0:r1]),
                f_nnz_global=n * args.features, input_dim=args.features, output_dim=args.classes)
    model = dict(hidden_dims=(args.hidden,), dropouts=(0.5, 0.5), lr=0.01, weight_decay=5e-4, seed=args.seed)
    comm = None
    t0 = time.perf_counter()
    if world > 1:
        comm = dmod.make_comm(eng, dist, rank, world, dev)
        g = eng.GCN(eng.PartDataset(part), comm=comm, **model)
    else:
        g = eng.GCN(eng.PartDataset(part), **model)
    torch.cuda.synchronize()
    t_create = time.perf_counter() - t0
    last = None
    for _ in range(args.warmup):
        last = (g.train_epoch(), g.eval(2))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    r = g.timed_epochs(args.steps, with_eval=True, time_graphsum=True)
    torch.cuda.synchronize()
    ms = torch.tensor([r["ms"], r["graphsum_ms"] / max(1, r["graphsum_calls"]), t_create * 1e3, t_gen * 1e3, t_graph * 1e3],
                      dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        free_b, total_b = torch.cuda.mem_get_info()
        line = {"workload": "scaleout_sym_community n=%d nnz=%d (community %d, intra %.0f, inter %.0f, sigma %.1f) f=%d c=%d; "
                            "2-layer GCN hidden %d; step = train_epoch + eval(2)" %
                            (n, nnz_global, args.block, args.intra, args.inter, args.sigma, args.features, args.classes, args.hidden),
                "n_gpus": world, "ms_per_epoch": float(ms[0]) / args.steps, "graphsum_mean_ms": float(ms[1]),
                "graphsum_calls_per_step": r["graphsum_calls"] / args.steps, "launches_per_step": r["launches"] / args.steps,
                "graph_staged": r.get("graph_staged"), "setup_ms": float(ms[2]), "gen_ms": float(ms[3]), "graph_gen_ms": float(ms[4]),
                "deg_max": int(deg_global.max()), "deg_mean": float(nnz_global / n), "train": last[0], "val": last[1],
                "gpu_mem_used_gb_rank0": (total_b - free_b) / 2**30, "steps": args.steps, "warmup": args.warmup,
                "host_cores": os.cpu_count()}
        print(json.dumps(line), flush=True)
    g.close()
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
