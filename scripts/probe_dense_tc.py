"""Exact-split tcgen05 GEMM (csrc/dense_tc.cu) against the fp32 SIMT kernels at the wide first layer's shape
(Reddit-shape: n = 232 965, f = 602, p = 600): parity against torch float64 on a row sample, then timing.
One JSON line per step appended to --out.  Run under `timeout`."""
import argparse
import json
import sys

import torch

sys.path.insert(0, ".")
import __graft_entry__ as ge

ge.load_package()
from parallel_gcn_b200 import binding as gcnb

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=232965)
ap.add_argument("--f", type=int, default=602)
ap.add_argument("--p", type=int, default=600)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--out", default="gpurun_out/probe_dense_tc.jsonl")
args = ap.parse_args()
dev = torch.device("cuda:0")
gcnb.device_check()
log = open(args.out, "a")


def emit(**kw):
    log.write(json.dumps(kw) + "\n")
    log.flush()
    print(json.dumps(kw), flush=True)


def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.iters


n, f, p = args.n, args.f, args.p
X = torch.randn(n, f, device=dev)
W = (torch.rand(f, p, device=dev) - 0.5) * 0.14
dH = torch.randn(n, p, device=dev) * 1e-3
img = gcnb.dense_tc_pack_x(X, n, f)
out = torch.full((n, p), float("nan"), device=dev)
gcnb.dense_tc_fwd(img, W, out, n, f, p)
torch.cuda.synchronize()
rows = torch.randint(0, n, (512,), device=dev)
want = X[rows].double() @ W.double()
err = (out[rows].double() - want).abs()
emit(step="fwd parity", max_err=float(err.max()), scale=float(want.abs().max()), nan=int(torch.isnan(out).sum()),
     bad=int((err > 1e-5 * want.abs() + 1e-6 * want.abs().max()).sum()))
ref = torch.empty(n, p, device=dev)
ms_tc = timeit(lambda: gcnb.dense_tc_fwd(img, W, out, n, f, p))
ms_simt = timeit(lambda: gcnb.matmul_nn(X, W, ref, n, f, p))
emit(step="fwd time", tcgen05_ms=ms_tc, simt_ms=ms_simt, tflops_fp32_equiv=2.0 * n * f * p / ms_tc / 1e9)

imgt = gcnb.dense_tc_pack_xt(X, n, f)
dW = torch.full((f, p), float("nan"), device=dev)
gcnb.dense_tc_tn(imgt, dH, dW, n, f, p)
torch.cuda.synchronize()
want = X.double().T @ dH.double()
err = (dW.double() - want).abs()
emit(step="tn parity", max_err=float(err.max()), scale=float(want.abs().max()), nan=int(torch.isnan(dW).sum()),
     bad=int((err > 1e-5 * want.abs() + 1e-6 * want.abs().max()).sum()))
dW2 = torch.empty(f, p, device=dev)
ms_tc = timeit(lambda: gcnb.dense_tc_tn(imgt, dH, dW, n, f, p))
ms_simt = timeit(lambda: gcnb.matmul_tn(X, dH, dW2, n, f, p))
emit(step="tn time", tcgen05_ms=ms_tc, simt_ms=ms_simt, tflops_fp32_equiv=2.0 * n * f * p / ms_tc / 1e9)
