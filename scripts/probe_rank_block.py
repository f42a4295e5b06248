"""What does ONE rank of an 8-way row-partitioned GraphSum pay for its local product?  ("~150 us of fixed
cost per GraphSum at 8 ranks".)  Single GPU, no communication: the bench graph's rows [r0, r0 + block) x all columns as a
device-built bit-tile plan, its steps timed apart (gcnb_bittile_debug_parts: 1 pack, 2 MMA kernel, 4 remainder) and together,
eager launches and one CUDA-graph replay of the same call.

  python scripts/probe_rank_block.py [--world 8] [--rank 3] [--iters 200]
"""
import argparse
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import __graft_entry__ as ge

ge.load_package()
import bench
from parallel_gcn_b200 import binding as gcnb
from parallel_gcn_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--rank", type=int, default=3)
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--scale", type=int, default=1)
args = ap.parse_args()
dev = torch.device("cuda:0")
gcnb.device_check()
w = bench.workload_config(args.scale)
indptr, indices = synth.synth_graph(w["n"], w["m"], n_blocks=w["blocks"], intra=w["intra"], sigma=w["sigma"], max_deg=w["max_deg"],
                                    seed=w["seed"])
n = len(indptr) - 1
deg = np.diff(indptr.astype(np.int64)).astype(np.float32)
s = (1.0 / np.sqrt(deg)).astype(np.float32)


def t_dev(a):
    a = np.ascontiguousarray(a)
    return torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).to(dev)


def timed(fn, iters):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


for world in sorted({1, args.world}):
    block = ((n + world - 1) // world + 3) // 4 * 4
    r0 = min(n, args.rank * block) if world > 1 else 0
    r1 = min(n, r0 + block)
    ip = (indptr[r0:r1 + 1] - indptr[r0]).astype(np.uint32)
    ix = indices[indptr[r0]:indptr[r1]]
    plan = gcnb.BitTilePlan.from_device(t_dev(ip), t_dev(ix), None, r1 - r0, n, t_dev(s[r0:r1].copy()), t_dev(s))
    assert plan is not None
    B = torch.randn(n, 16, device=dev)
    C = torch.empty(r1 - r0, 16, device=dev)
    rec = {"world": world, "rows": r1 - r0, "entries": int(len(ix)), "info": plan.info()}
    for name, mask in (("pack", 1), ("mma", 2), ("remainder", 4), ("all", 15)):
        plan.debug_parts(mask)
        rec[name + "_us"] = round(timed(lambda: plan.spmm16(B, C), args.iters), 2)
    plan.debug_parts(15)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        plan.spmm16(B, C)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            plan.spmm16(B, C)
        rec["all_graph_replay_us"] = round(timed(g.replay, args.iters), 2)
    print(json.dumps(rec), flush=True)
    plan.close()
