"""Locality reordering on the GPU: the Reddit-shape graph with its node ids SHUFFLED (how a file usually
arrives) -> GraphSum d=16 falls back to the generic kernel; gcnb_reorder_communities (label propagation, host) renumbers
it -> the window-staged kernels apply again.  Prints one JSON line."""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.load_package()
gcnb = importlib.import_module("parallel_gcn_b200.binding")
eng = importlib.import_module("parallel_gcn_b200.engine")
import bench  # noqa: E402

dev = torch.device("cuda:0")
gcnb.device_check()
w = bench.workload_config(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
ip, ix = eng.synth_graph(w["n"], w["m"], w["blocks"], w["intra"], w["sigma"], w["max_deg"], w["seed"])
n = w["n"]


def permute(ip_, ix_, perm):
    op, ox = np.empty(n + 1, np.uint32), np.empty(len(ix_), np.uint32)
    eng.check(eng.lib.gcnb_permute_csr(n, eng._p(ip_), eng._p(ix_), eng._p(perm), eng._p(op), eng._p(ox)))
    return op, ox


def graphsum_us(ip_, ix_, iters=10):
    d_ip, d_ix = torch.from_numpy(ip_.view(np.int32)).to(dev), torch.from_numpy(ix_.view(np.int32)).to(dev)
    vals = torch.empty(len(ix_), device=dev)
    gcnb.check(gcnb.lib.gcnb_graph_values_f32(gcnb.ptr(d_ip), gcnb.ptr(d_ix), n, gcnb.ptr(vals), gcnb.stream()))
    plan = gcnb.SpmmPlan(d_ip, d_ix, n)
    info = plan.stage(vals, 16, ip_, ix_)
    x, out = torch.randn(n, 16, device=dev), torch.empty(n, 16, device=dev)
    for _ in range(3):
        plan.spmm(vals, x, out, 16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        plan.spmm(vals, x, out, 16)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    plan.close()
    return us, info["staged"], info["staged_nnz"] / len(ix_)


res = {"workload": "reddit_shape n=%d nnz=%d" % (n, len(ix))}
res["community_order"] = graphsum_us(ip, ix)
shuffle = np.random.default_rng(0).permutation(n).astype(np.uint32)
sp, sx = permute(ip, ix, shuffle)
res["shuffled_ids"] = graphsum_us(sp, sx)
t0 = time.perf_counter()
perm, n_comm = eng.reorder_communities(sp, sx)
res["reorder_s"] = time.perf_counter() - t0
res["communities_found"] = n_comm
rp, rx = permute(sp, sx, perm)
res["reordered"] = graphsum_us(rp, rx)
res["columns"] = "graphsum_us, staged?, staged fraction of the entries"
res["host_cores"] = os.cpu_count()
print(json.dumps(res), flush=True)
