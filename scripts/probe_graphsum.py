"""Quick GraphSum probe on a Reddit-shape random CSR (device-generated; perf only, not parity)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import __graft_entry__ as ge

ge.load_package()
from parallel_gcn_b200 import binding as gcnb

dev = torch.device("cuda:0")
gcnb.device_check()
N, MEAN, BLOCKS = 232965, 492, 50


def make(intra, seed=0, sort=True):
    g = torch.Generator(device=dev).manual_seed(seed)
    deg = torch.exp(torch.randn(N, device=dev, generator=g) * 1.2)
    deg = (deg * (MEAN / deg.mean())).clamp(1, 21657)
    deg = (deg * (MEAN / deg.mean())).clamp(1, 21657).round().long()
    indptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(deg, 0)
    nnz = int(indptr[-1])
    rows = torch.repeat_interleave(torch.arange(N, device=dev), deg)
    bs = (N + BLOCKS - 1) // BLOCKS
    local = (rows // bs) * bs + torch.randint(0, bs, (nnz,), device=dev, generator=g)
    local = local.clamp(max=N - 1)
    glob = torch.randint(0, N, (nnz,), device=dev, generator=g)
    pick = torch.rand(nnz, device=dev, generator=g) < intra
    cols = torch.where(pick, local, glob)
    if sort:
        key = rows * N + cols
        key, _ = torch.sort(key)
        cols = key % N
    vals = torch.rand(nnz, device=dev, generator=g)
    return indptr.int(), cols.int(), vals, nnz


def bench(plan, vals, x, out, dim, iters=20):
    for _ in range(3):
        plan.spmm(vals, x, out, dim)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        plan.spmm(vals, x, out, dim)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


import argparse
ap = argparse.ArgumentParser()
ap.add_argument("--intra", type=float, nargs="*", default=[0.8, 0.0])
ap.add_argument("--dim", type=int, nargs="*", default=[16, 41, 64])
ap.add_argument("--seg", type=int, nargs="*", default=[256, 512, 1024])
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--stage", type=int, default=1)
ap.add_argument("--min-seg", type=int, nargs="*", default=[0])
ap.add_argument("--seg-cap", type=int, nargs="*", default=[0])
ap.add_argument("--rows-div", type=int, nargs="*", default=[1], help="keep only the first N/div rows (a rank's row block)")
ap.add_argument("--out", default="gpurun_out/probe_graphsum.json")
args = ap.parse_args()
res = []
for intra, div in [(a, b) for a in args.intra for b in args.rows_div]:
    indptr, cols, vals, nnz = make(intra)
    nrows = N // div
    if div > 1:
        nnz = int(indptr[nrows])
        indptr, cols, vals = indptr[: nrows + 1].contiguous(), cols[:nnz].contiguous(), vals[:nnz].contiguous()
    for dim in args.dim:
        x = torch.randn(N, dim, device=dev)
        out = torch.empty(nrows, dim, device=dev)
        for seg in args.seg:
          for min_seg in args.min_seg:
            for seg_cap in args.seg_cap:
                plan = gcnb.SpmmPlan(indptr, cols, N, seg)
                us0 = bench(plan, vals, x, out, dim, args.iters)
                ref = out.clone()
                sinfo, build_s = None, None
                if args.stage and dim == 16:
                    t0 = time.time()
                    sinfo = plan.stage(vals, dim, min_seg=min_seg, seg_cap=seg_cap)
                    build_s = round(time.time() - t0, 2)
                us = bench(plan, vals, x, out, dim, args.iters)
                err = float((out - ref).abs().max() / ref.abs().max())
                alg = 4 * (nrows + 1) + 8 * nnz + 4 * (N + nrows) * dim
                r = dict(rows_div=div, intra=intra, dim=dim, seg=seg, min_seg=min_seg, seg_cap=seg_cap, nnz=nnz, us_generic=round(us0, 1),
                         us=round(us, 1), alg_GBs=round(alg / us / 1e3, 1), rel_diff_vs_generic=err, stage=sinfo,
                         stage_build_s=build_s, info=plan.info())
                print(json.dumps(r), flush=True)
                res.append(r)
                plan.close()
json.dump(res, open(args.out, "w"), indent=1)
