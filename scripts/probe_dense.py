"""Times the first-layer kernels at Reddit shape (N=232965, F=602, P=16)."""
import sys
import torch
sys.path.insert(0, ".")
import __graft_entry__ as ge
ge.load_package()
from parallel_gcn_b200 import binding as gcnb
dev = torch.device("cuda:0")
N, F, P = 232965, 602, 16
X = torch.randn(N, F, device=dev)
W = torch.randn(F, P, device=dev)
dH = torch.randn(N, P, device=dev)
out = torch.empty(N, P, device=dev)
dW = torch.empty(F, P, device=dev)
Xd = torch.empty_like(X)
bits = torch.zeros(gcnb.lib.gcnb_dropout_maskbits_words(N, F), dtype=torch.int32, device=dev)
rng = gcnb.make_rng(1, [(F * P, 1)])
ws_tn = torch.empty((gcnb.lib.gcnb_dense_feat_tn_workspace(N, F, P) + 3) // 4, device=dev)
ws_old = torch.empty((gcnb.lib.gcnb_matmul_tn_workspace(N, F, P) + 3) // 4, device=dev)
import ctypes as C
def t(name, fn, iters=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print("%-28s %9.1f us   (%.0f GB/s of X traffic)" % (name, us, N * F * 4 / us / 1e3), flush=True)
t("maskbits", lambda: gcnb.dropout_maskbits(bits, N, F, 0.5, rng))
t("dense_feat_fwd (bits)", lambda: gcnb.dense_feat_fwd(X, bits, 0.5, W, out, N, F, P))
t("dense_feat_fwd (no bits)", lambda: gcnb.dense_feat_fwd(X, None, 0.0, W, out, N, F, P))
t("dense_feat_tn (bits)", lambda: gcnb.dense_feat_tn(X, bits, 0.5, dH, dW, N, F, P, ws_tn))
t("dense_feat_tn (no bits)", lambda: gcnb.dense_feat_tn(X, None, 0.0, dH, dW, N, F, P, ws_tn))
t("old dropout_oop", lambda: gcnb.check(gcnb.lib.gcnb_dropout_fwd_oop_f32(gcnb.ptr(X), gcnb.ptr(Xd), None, None, N * F, 0.5, C.byref(rng), gcnb.stream())))
t("old matmul_nn", lambda: gcnb.matmul_nn(X, W, out, N, F, P))
t("old matmul_tn", lambda: gcnb.matmul_tn(X, dH, dW, N, F, P, ws_old))
lg = torch.randn(N, 41, device=dev); gr = torch.empty_like(lg)
truth = torch.randint(-1, 41, (N,), device=dev, dtype=torch.int32)
res = torch.zeros(4, device=dev); wsce = gcnb.zeroed_workspace(gcnb.lib.gcnb_ce_workspace(N), dev)
t("softmax_ce train", lambda: gcnb.softmax_ce(lg, gr, truth, N, 41, 150000, True, res, wsce))
