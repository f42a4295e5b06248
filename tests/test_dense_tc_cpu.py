"""Exact-split GEMM (parallel-gcn_b200/csrc/dense_tc.cu), the arithmetic checked on the CPU: x and w as three truncated
bf16 pieces each, the six piece products with (piece_x + piece_w) <= 2 accumulated exactly -- within fp32 rounding of the
float64 product at the wide layer's shape (f = 602).  The kernel itself has an opt-in GPU test (tests/test_zz_bittile_gpu.py)."""
import numpy as np

from tests.util import assert_close


def split3(x):
    x = x.astype(np.float32)
    h = (x.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)
    r1 = x - h
    m = (r1.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)
    r2 = r1 - m
    lo = (r2.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)
    assert np.array_equal(h.astype(np.float64) + m + lo, x.astype(np.float64))
    return [p.astype(np.float64) for p in (h, m, lo)]


def test_six_piece_products_reproduce_the_fp32_product():
    rng = np.random.default_rng(0)
    n, f, p = 256, 602, 600
    X = rng.standard_normal((n, f)).astype(np.float32)
    W = ((rng.random((f, p)) - 0.5) * 2 * np.sqrt(6.0 / (f + p))).astype(np.float32)  # Glorot range
    want = X.astype(np.float64) @ W.astype(np.float64)
    xs, ws = split3(X), split3(W)
    got = np.zeros_like(want)
    for i, j in ((2, 0), (0, 2), (1, 1), (1, 0), (0, 1), (0, 0)):  # the kernel's order: smallest terms first
        got += xs[i] @ ws[j]
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 2e-7 * scale
    assert_close(got, want, what="six-term split product")
    # the three dropped products together are below 2^-22 of the largest result; four terms would not be enough
    four = xs[0] @ ws[0] + xs[0] @ ws[1] + xs[1] @ ws[0] + xs[1] @ ws[1]
    assert np.abs(four - want).max() > np.abs(got - want).max() * 10


def test_shape_helpers(monkeypatch):
    import importlib
    import __graft_entry__ as ge
    ge.load_package()
    gcnb = importlib.import_module("parallel_gcn_b200.binding")
    assert gcnb.lib.gcnb_dense_tc_supported(602, 600) == 1 and gcnb.lib.gcnb_dense_tc_supported(602, 8) == 0
    n_blk, ks = (232965 + 127) // 128, (602 + 15) // 16
    assert gcnb.lib.gcnb_dense_tc_x_bytes(232965, 602) == n_blk * ks * 3 * 4096
    assert gcnb.lib.gcnb_dense_tc_w_bytes(602, 600) == 3 * ks * 3 * 208 * 32  # 600 columns -> 3 parts of 208
    assert gcnb.lib.gcnb_dense_tc_w_bytes(50, 16) == 1 * 4 * 3 * 16 * 32
