"""GPU parity of every C-ABI kernel against the oracle (gcn_oracle.c) on seeded inputs.

Tolerances: integer / mask / index results bit-exact; fp32 results within 1e-5 relative (north_star), with an
absolute floor of 1e-6 x max|want| where sums cancel.  All calls go through the C ABI (ctypes)."""
import ctypes as C

import numpy as np
import pytest

from tests.util import assert_close, random_csr, to_dev, to_np

pytestmark = pytest.mark.gpu
f32, u32, i32, u8 = np.float32, np.uint32, np.int32, np.uint8


def _graphsum_check(O, gcnb, dev, indptr, indices, values, dim, seg_nnz=0, n_cols=None):
    import torch
    n = len(indptr) - 1
    n_cols = n if n_cols is None else n_cols
    rng = np.random.default_rng(dim * 1000 + n)
    x = rng.standard_normal((n_cols, dim)).astype(f32)
    want = np.empty((n, dim), f32)
    O.lib.orc_spmm(n, dim, O._p(indptr), O._p(indices), O._p(values), O._p(x), O._p(want))
    d_ip, d_ix, d_v, d_x = (to_dev(a, dev) for a in (indptr, indices, values, x))
    out = torch.full((n, dim), float("nan"), device=dev)
    plan = gcnb.SpmmPlan(d_ip, d_ix, n_cols, seg_nnz)
    plan.spmm(d_v, d_x, out, dim)
    torch.cuda.synchronize()
    assert_close(to_np(out), want, what="spmm dim=%d" % dim)
    # determinism: bit-identical on a second launch
    out2 = torch.empty_like(out)
    plan.spmm(d_v, d_x, out2, dim)
    assert torch.equal(out, out2)
    info = plan.info()
    plan.close()
    return info


@pytest.mark.parametrize("name", ["cora", "citeseer"])
@pytest.mark.parametrize("dim", [16, 7, 41, 3, 64, 600])
def test_graphsum_datasets(O, gcnb, dev, datasets, name, dim):
    ds = datasets[name]
    _graphsum_check(O, gcnb, dev, ds.g_indptr, ds.g_indices, ds.graph_values(), dim)


@pytest.mark.parametrize("dim", [16, 41, 128, 5])
def test_graphsum_skewed_split_rows(O, gcnb, dev, dim):
    """rows far longer than a segment (multi-segment combine), empty rows, tiny segments."""
    rng = np.random.default_rng(7)
    n = 3000
    indptr, indices = random_csr(rng, n, n, 12, heavy_rows=[(5, 20000), (6, 513), (2999, 7777)], empty_rows=[0, 17, 2998])
    values = rng.standard_normal(len(indices)).astype(f32)
    info = _graphsum_check(O, gcnb, dev, indptr, indices, values, dim, seg_nnz=64)
    assert info["n_split_rows"] >= 3 and info["max_deg"] == 20000


@pytest.mark.parametrize("cfg", [
    dict(n=3000, comm=6, deg=60, intra=0.8, window=512, min_seg=8, seg_cap=64, heavy=((5, 2500), (2999, 900)), h=True),
    dict(n=1000, comm=2, deg=30, intra=0.9, window=500, min_seg=4, seg_cap=128, heavy=(), h=False),
    dict(n=20000, comm=5, deg=100, intra=0.8, window=0, min_seg=0, seg_cap=0, heavy=((77, 2500),), h=False),
])
def test_graphsum_window_staged(O, gcnb, dev, cfg):
    """window-staged GraphSum (spmm_stage.cu): shared-memory gathers for the clustered part, generic kernel for the
    remainder, partial rows added in slot order -- same product as the oracle, bit-identical between launches, and
    untouched behaviour for calls the staging does not cover (other value array, other dim)."""
    import torch
    from tests.test_stage_cpu import community_csr
    rng = np.random.default_rng(21)
    n, dim = cfg["n"], 16
    indptr, indices = community_csr(rng, n, cfg["comm"], cfg["deg"], cfg["intra"], cfg["heavy"])
    values = rng.standard_normal(len(indices)).astype(f32)
    x = rng.standard_normal((n, dim)).astype(f32)
    want = np.empty((n, dim), f32)
    O.lib.orc_spmm(n, dim, O._p(indptr), O._p(indices), O._p(values), O._p(x), O._p(want))
    d_ip, d_ix, d_v, d_x = (to_dev(a, dev) for a in (indptr, indices, values, x))
    plan = gcnb.SpmmPlan(d_ip, d_ix, n)
    base = torch.empty((n, dim), device=dev)
    plan.spmm(d_v, d_x, base, dim)
    info = plan.stage(d_v, dim, indptr if cfg["h"] else None, indices if cfg["h"] else None, cfg["window"],
                      cfg["min_seg"], cfg["seg_cap"], 1 if cfg["window"] else 0)
    assert info["staged"] == 1 and info["staged_nnz"] + info["rem_nnz"] == len(indices)
    assert info["staged_nnz"] > 0.5 * len(indices)
    out = torch.full((n, dim), float("nan"), device=dev)
    plan.spmm(d_v, d_x, out, dim)
    torch.cuda.synchronize()
    assert_close(to_np(out), want, what="staged spmm")
    out2 = torch.full((n, dim), float("nan"), device=dev)
    plan.spmm(d_v, d_x, out2, dim)
    assert torch.equal(out, out2)
    # a different value array / dim falls through to the generic kernel and is unaffected
    d_v2 = d_v.clone()
    out3 = torch.empty_like(out)
    plan.spmm(d_v2, d_x, out3, dim)
    assert torch.equal(out3, base)
    x8 = torch.randn(n, 8, device=dev)
    o8, o8b = torch.empty(n, 8, device=dev), torch.empty(n, 8, device=dev)
    plan.spmm(d_v, x8, o8, 8)
    plan2 = gcnb.SpmmPlan(d_ip, d_ix, n)
    plan2.spmm(d_v, x8, o8b, 8)
    assert torch.equal(o8, o8b)
    # new values: re-gather only
    d_v3 = d_v * 0.5
    plan.stage(d_v3, dim)
    plan.spmm(d_v3, d_x, out2, dim)
    torch.cuda.synchronize()
    assert_close(to_np(out2), want * 0.5, what="staged spmm after value change")
    plan.close()
    plan2.close()


@pytest.mark.parametrize("dim,ldb,ldc,b_off,c_off", [(16, 40, 24, 8, 4), (7, 9, 11, 2, 3), (64, 64, 96, 0, 32), (10, 602, 602, 592, 592)])
def test_graphsum_column_slabs_generic(O, gcnb, dev, dim, ldb, ldc, b_off, c_off):
    """gcnb_spmm_ld_f32: B and C are column slabs of wider matrices; columns outside the slab are neither read nor
    written."""
    import torch
    rng = np.random.default_rng(5)
    n = 2500
    indptr, indices = random_csr(rng, n, n, 20, heavy_rows=[(9, 5000)], empty_rows=[0, 1234])
    values = rng.standard_normal(len(indices)).astype(f32)
    xw = rng.standard_normal((n, ldb)).astype(f32)
    x = np.ascontiguousarray(xw[:, b_off:b_off + dim])
    want = np.empty((n, dim), f32)
    O.lib.orc_spmm(n, dim, O._p(indptr), O._p(indices), O._p(values), O._p(x), O._p(want))
    d_ip, d_ix, d_v, d_x = (to_dev(a, dev) for a in (indptr, indices, values, xw))
    plan = gcnb.SpmmPlan(d_ip, d_ix, n, 64)
    out = torch.full((n, ldc), -7.0, device=dev)
    plan.spmm_ld(d_v, d_x, ldb, out, ldc, dim, b_off=b_off, c_off=c_off)
    torch.cuda.synchronize()
    got = to_np(out)
    assert_close(got[:, c_off:c_off + dim], want, what="slab spmm")
    mask = np.ones(ldc, bool)
    mask[c_off:c_off + dim] = False
    assert (got[:, mask] == -7.0).all()
    plan.close()


@pytest.mark.parametrize("dim,ld", [(64, 64), (600, 600), (41, 41), (602, 602), (32, 50)])
def test_graphsum_window_staged_wide(O, gcnb, dev, dim, ld, monkeypatch):
    """operands wider than the staged width run as 16-column slabs through the staged kernels (last slab shifted left to
    end at dim): same product as the oracle, other columns of a wider C untouched, bit-identical between launches."""
    import torch
    from tests.test_stage_cpu import community_csr
    rng = np.random.default_rng(33)
    n = 6000
    indptr, indices = community_csr(rng, n, 4, 80, 0.8, ((5, 2500),))
    values = rng.standard_normal(len(indices)).astype(f32)
    xw = rng.standard_normal((n, ld)).astype(f32)
    x = np.ascontiguousarray(xw[:, :dim])
    want = np.empty((n, dim), f32)
    O.lib.orc_spmm(n, dim, O._p(indptr), O._p(indices), O._p(values), O._p(x), O._p(want))
    d_ip, d_ix, d_v, d_x = (to_dev(a, dev) for a in (indptr, indices, values, xw))
    plan = gcnb.SpmmPlan(d_ip, d_ix, n)
    info = plan.stage(d_v, 16, None, None, 1024, 8, 64, 1)
    assert info["staged"] == 1
    out = torch.full((n, ld), -7.0, device=dev)
    plan.spmm_ld(d_v, d_x, ld, out, ld, dim)
    torch.cuda.synchronize()
    got = to_np(out)
    assert_close(got[:, :dim], want, what="staged slab spmm dim=%d" % dim)
    assert (got[:, dim:] == -7.0).all()
    out2 = torch.full((n, ld), -7.0, device=dev)
    plan.spmm_ld(d_v, d_x, ld, out2, ld, dim)
    assert torch.equal(out, out2)
    plan.close()


def test_graphsum_window_staged_own_slab_first(O, gcnb, dev):
    """row-partitioned GraphSum: the staged windows inside the rank's own column range are launched from the own slab
    before the gathered matrix exists (gcnb_spmm_stage_own_f32), the rest afterwards -- same product, bit-identical to
    the one-call form."""
    import torch
    from tests.test_stage_cpu import community_csr
    rng = np.random.default_rng(17)
    n_rows, n_cols, dim = 2000, 6000, 16   # a row block [2000, 4000) of a 6000-node graph
    own = (2000, 4000)
    ip, ix = community_csr(rng, n_cols, 6, 60, 0.8)
    indptr = (ip[own[0]:own[1] + 1] - ip[own[0]]).astype(u32)
    indices = np.ascontiguousarray(ix[ip[own[0]]:ip[own[1]]])
    values = rng.standard_normal(len(indices)).astype(f32)
    x = rng.standard_normal((n_cols, dim)).astype(f32)
    want = np.empty((n_rows, dim), f32)
    O.lib.orc_spmm(n_rows, dim, O._p(indptr), O._p(indices), O._p(values), O._p(x), O._p(want))
    d_ip, d_ix, d_v, d_x = (to_dev(a, dev) for a in (indptr, indices, values, x))
    plan = gcnb.SpmmPlan(d_ip, d_ix, n_cols)
    plan.set_own_cols(*own)
    info = plan.stage(d_v, dim, None, None, 512, 8, 64, 1)
    assert info["staged"] == 1
    one = torch.full((n_rows, dim), float("nan"), device=dev)
    plan.spmm(d_v, d_x, one, dim)            # everything from the full matrix
    torch.cuda.synchronize()
    assert_close(to_np(one), want, what="staged spmm, own + remote runs in one call")
    own_slab = d_x[own[0]:own[1]].clone()    # only the rank's rows exist at this point
    full_late = torch.full_like(d_x, float("nan"))
    assert plan.stage_own(d_v, own_slab, dim)
    full_late.copy_(d_x)                     # "the exchange completes"
    two = torch.full((n_rows, dim), float("nan"), device=dev)
    plan.spmm(d_v, full_late, two, dim)
    torch.cuda.synchronize()
    assert torch.equal(one, two)
    plan.close()


@pytest.mark.parametrize("name", ["cora", "citeseer", "synthetic"])
def test_graph_values_on_device_bit_exact(O, gcnb, dev, datasets, name):
    """Parser::calculateGraphValues on the device equals the parser's host values bit for bit (sqrtf and the double
    divide are IEEE on both sides)."""
    import torch
    if name == "synthetic":
        rng = np.random.default_rng(2)
        n = 5000
        indptr, indices = random_csr(rng, n, n, 30, heavy_rows=[(3, 4000), (77, 65000 // 8)], empty_rows=[])
        deg = np.diff(indptr.astype(np.int64)).astype(np.uint32)
        rows = np.repeat(np.arange(n), deg)
        prod = (deg[rows] * deg[indices]).astype(f32)
        want = (1.0 / np.sqrt(prod, dtype=f32).astype(np.float64)).astype(f32)
    else:
        ds = datasets[name]
        indptr, indices, want = ds.g_indptr, ds.g_indices, ds.graph_values()
    d_ip, d_ix = to_dev(indptr, dev), to_dev(indices, dev)
    out = torch.empty(len(indices), device=dev)
    gcnb.check(gcnb.lib.gcnb_graph_values_f32(gcnb.ptr(d_ip), gcnb.ptr(d_ix), len(indptr) - 1, gcnb.ptr(out), gcnb.stream()))
    torch.cuda.synchronize()
    assert np.array_equal(to_np(out).view(u32), np.asarray(want, f32).view(u32))


def test_graphsum_ref_cpu_flavour(O, gcnb, dev, datasets):
    """the ref-CPU GraphSum recomputes coef per edge (module.cpp:86-90); hoisted values give the same bits."""
    ds = datasets["cora"]
    n, dim = ds.num_nodes, 16
    x = np.random.default_rng(0).standard_normal((n, dim)).astype(f32)
    a, b = np.empty((n, dim), f32), np.empty((n, dim), f32)
    O.lib.orc_graphsum(n, dim, O._p(ds.g_indptr), O._p(ds.g_indices), None, O._p(x), O._p(a))
    O.lib.orc_graphsum(n, dim, O._p(ds.g_indptr), O._p(ds.g_indices), O._p(ds.graph_values()), O._p(x), O._p(b))
    assert (a == b).all()


@pytest.mark.parametrize("name,p", [("cora", 16), ("citeseer", 16), ("cora", 72)])
def test_sparse_matmul_fwd_bwd(O, gcnb, dev, datasets, name, p):
    import torch
    ds = datasets[name]
    n, F = ds.num_nodes, ds.input_dim
    rng = np.random.default_rng(3)
    w = rng.standard_normal((F, p)).astype(f32)
    vals = (ds.f_value * rng.integers(0, 2, len(ds.f_value)) * 2).astype(f32)  # as after dropout
    cg = rng.standard_normal((n, p)).astype(f32)
    want_c, want_g = np.empty((n, p), f32), np.empty((F, p), f32)
    O.lib.orc_spmm(n, p, O._p(ds.f_indptr), O._p(ds.f_indices), O._p(vals), O._p(w), O._p(want_c))
    O.lib.orc_spmm_bwd(n, F, p, O._p(ds.f_indptr), O._p(ds.f_indices), O._p(vals), O._p(cg), O._p(want_g))
    d_ip, d_ix, d_v, d_w, d_cg = (to_dev(a, dev) for a in (ds.f_indptr, ds.f_indices, vals, w, cg))
    plan = gcnb.SpmmPlan(d_ip, d_ix, F)
    c = torch.empty((n, p), device=dev)
    plan.spmm(d_v, d_w, c, p)
    assert_close(to_np(c), want_c, what="sparse_matmul fwd")
    csc = gcnb.Csc(d_ip, d_ix, F)
    assert not csc.is_dense
    g = torch.full((F, p), float("nan"), device=dev)
    csc.plan.spmm(d_v, d_cg, g, p, perm=csc.perm)
    torch.cuda.synchronize()
    assert_close(to_np(g), want_g, what="sparse_matmul bwd")
    # CSC bit-exactness against a host transpose
    colptr = to_np(csc.colptr, u32)
    rowidx = to_np(csc.rowidx, u32)
    order = np.argsort(ds.f_indices, kind="stable")
    rows = np.repeat(np.arange(n, dtype=u32), np.diff(ds.f_indptr.astype(np.int64)))
    assert (rowidx == rows[order]).all() and (to_np(csc.perm, u32) == order.astype(u32)).all()
    assert (colptr == np.concatenate([[0], np.cumsum(np.bincount(ds.f_indices, minlength=F))]).astype(u32)).all()
    csc.close(); plan.close()


def test_csc_dense_detection(gcnb, dev):
    n, F = 37, 12
    indptr = (np.arange(n + 1) * F).astype(u32)
    indices = np.tile(np.arange(F, dtype=u32), n)
    csc = gcnb.Csc(to_dev(indptr, dev), to_dev(indices, dev), F)
    assert csc.is_dense
    csc.close()


@pytest.mark.parametrize("m,n,p", [(2708, 16, 7), (3327, 16, 6), (5000, 16, 41), (1000, 72, 7), (777, 602, 16),
                                   (513, 600, 41), (300, 130, 70), (1, 16, 41), (129, 1, 1), (1500, 602, 600), (260, 64, 33),
                                   (20011, 16, 41), (9000, 30, 64), (4097, 7, 3)])
def test_matmul_nn_nt_tn(O, gcnb, dev, m, n, p):
    import torch
    rng = np.random.default_rng(m + n + p)
    a = rng.standard_normal((m, n)).astype(f32)
    b = rng.standard_normal((n, p)).astype(f32)
    cg = rng.standard_normal((m, p)).astype(f32)
    want_c, want_ag, want_bg = np.empty((m, p), f32), np.empty((m, n), f32), np.empty((n, p), f32)
    O.lib.orc_matmul(m, n, p, O._p(a), O._p(b), O._p(want_c))
    O.lib.orc_matmul_bwd(m, n, p, O._p(a), O._p(b), O._p(cg), O._p(want_ag), O._p(want_bg))
    d_a, d_b, d_cg = (to_dev(x, dev) for x in (a, b, cg))
    c = torch.full((m, p), float("nan"), device=dev)
    ag = torch.full((m, n), float("nan"), device=dev)
    bg = torch.full((n, p), float("nan"), device=dev)
    gcnb.matmul_nn(d_a, d_b, c, m, n, p)
    gcnb.matmul_nt(d_cg, d_b, ag, m, n, p)
    gcnb.matmul_tn(d_a, d_cg, bg, m, n, p)
    torch.cuda.synchronize()
    assert_close(to_np(c), want_c, what="matmul nn")
    assert_close(to_np(ag), want_ag, what="matmul nt")
    assert_close(to_np(bg), want_bg, rtol=2e-5, what="matmul tn")
    bg2 = torch.empty_like(bg)
    gcnb.matmul_tn(d_a, d_cg, bg2, m, n, p)
    assert torch.equal(bg, bg2), "weight gradient must be deterministic"


def _oracle_draws(history, size):
    g = (size + 3) // 4
    d = np.zeros(g, u32)
    for s, c in history:
        d[: min(g, (s + 3) // 4)] += c
    return d


@pytest.mark.parametrize("size,rows,cols,hist", [(1433 * 16, 1433, 16, []), (16 * 7, 16, 7, [(1433 * 16, 1)]),
                                                 (10, 3, 3, [(5, 2), (100, 1)])])
def test_glorot_bit_exact(O, gcnb, dev, size, rows, cols, hist):
    import torch
    seed = 19990304
    want = np.empty(size, f32)
    O.lib.orc_glorot_philox(size, rows, cols, seed, O._p(_oracle_draws(hist, size)), 1, O._p(want))
    w = torch.empty(size, device=dev)
    gcnb.glorot(w, rows, cols, gcnb.make_rng(seed, hist))
    assert (to_np(w).view(u32) == want.view(u32)).all()


@pytest.mark.parametrize("size,p,hist", [(49216, 0.5, [(22928, 1), (112, 1)]), (43328, 0.5, [(49216, 3), (22928, 1), (112, 1), (43328, 2)]),
                                         (1001, 0.6, []), (7, 0.0, []), (4096, 0.1, [(10, 1)])])
def test_dropout_philox_bit_exact(O, gcnb, dev, size, p, hist):
    import torch
    seed = 311288059
    rng = np.random.default_rng(size)
    x = rng.standard_normal(size).astype(f32)
    mask_want = np.empty(size, u8)
    O.lib.orc_dropout_mask_philox(size, p, seed, O._p(_oracle_draws(hist, size)), 1, O._p(mask_want))
    want = x.copy()
    O.lib.orc_dropout_apply(size, O._p(want), O._p(mask_want), O.lib.orc_dropout_scale(p, 1))
    d_x = to_dev(x, dev)
    d_m = torch.zeros(size, dtype=torch.uint8, device=dev)
    gcnb.dropout_fwd(d_x, d_m, p, rng=gcnb.make_rng(seed, hist))
    assert (to_np(d_m) == mask_want).all()
    assert (to_np(d_x).view(u32) == want.view(u32)).all()
    # no-mask variant (input Variable has no grad) and injected masks
    d_x2 = to_dev(x, dev)
    gcnb.dropout_fwd(d_x2, None, p, rng=gcnb.make_rng(seed, hist))
    assert torch.equal(d_x, d_x2)
    d_x3 = to_dev(x, dev)
    gcnb.dropout_fwd(d_x3, None, p, ext_mask=to_dev(mask_want, dev))
    assert torch.equal(d_x, d_x3)
    # backward
    g = rng.standard_normal(size).astype(f32)
    gw = g.copy()
    O.lib.orc_dropout_apply(size, O._p(gw), O._p(mask_want), O.lib.orc_dropout_scale(p, 1))
    d_g = to_dev(g, dev)
    gcnb.dropout_bwd(d_g, d_m, p)
    assert (to_np(d_g).view(u32) == gw.view(u32)).all()


@pytest.mark.parametrize("size", [1, 43328, 100003])
def test_relu_and_fused_relu_dropout(O, gcnb, dev, size):
    import torch
    rng = np.random.default_rng(size)
    x = rng.standard_normal(size).astype(f32)
    x[::7] = 0.0
    g = rng.standard_normal(size).astype(f32)
    p, seed = 0.5, 123
    xw, rm = x.copy(), np.zeros(size, u8)
    O.lib.orc_relu_fwd(size, O._p(xw), O._p(rm), 1)
    d_x, d_m = to_dev(x, dev), torch.zeros(size, dtype=torch.uint8, device=dev)
    gcnb.relu_fwd(d_x, d_m, True)
    assert (to_np(d_x) == xw).all() and (to_np(d_m) == rm).all()
    gw = g.copy(); O.lib.orc_relu_bwd(size, O._p(gw), O._p(rm))
    d_g = to_dev(g, dev); gcnb.relu_bwd(d_g, d_m)
    assert (to_np(d_g) == gw).all()
    # eval mode leaves the mask untouched
    d_m2 = torch.full((size,), 9, dtype=torch.uint8, device=dev)
    gcnb.relu_fwd(to_dev(x, dev), d_m2, False)
    assert (to_np(d_m2) == 9).all()
    # fused ReLU+Dropout == the two modules chained
    dm = np.empty(size, u8)
    O.lib.orc_dropout_mask_philox(size, p, seed, None, 1, O._p(dm))
    O.lib.orc_dropout_apply(size, O._p(xw), O._p(dm), 2.0)
    d_x, d_mm = to_dev(x, dev), torch.zeros(size, dtype=torch.uint8, device=dev)
    gcnb.relu_dropout_fwd(d_x, d_mm, p, True, rng=gcnb.make_rng(seed))
    assert (to_np(d_x).view(u32) == xw.view(u32)).all()
    assert (to_np(d_mm) == (rm | (dm << 1))).all()
    gw = g.copy(); O.lib.orc_dropout_apply(size, O._p(gw), O._p(dm), 2.0); O.lib.orc_relu_bwd(size, O._p(gw), O._p(rm))
    d_g = to_dev(g, dev); gcnb.relu_dropout_bwd(d_g, d_mm, p)
    assert (to_np(d_g).view(u32) == gw.view(u32)).all()
    # fused eval: relu only
    d_x = to_dev(x, dev); gcnb.relu_dropout_fwd(d_x, None, p, False)
    xe = x.copy(); O.lib.orc_relu_fwd(size, O._p(xe), O._p(rm), 0)
    assert (to_np(d_x) == xe).all()


@pytest.mark.parametrize("n,C,training", [(2708, 7, 1), (3327, 6, 1), (5000, 41, 1), (5000, 41, 0), (100, 172, 1), (3, 2, 1)])
def test_softmax_ce(O, gcnb, dev, n, C, training):
    import torch
    rng = np.random.default_rng(n * C)
    logits = (rng.standard_normal((n, C)) * 3).astype(f32)
    truth = rng.integers(-1, C, n).astype(i32)
    truth[rng.random(n) < 0.4] = -1
    ns = int((truth >= 0).sum()) + 5  # split count may exceed labelled rows (citeseer)
    lw, gw = logits.copy(), np.empty((n, C), f32)
    cnt = np.zeros(1, np.int64)
    loss_want = O.lib.orc_cross_entropy(n, C, O._p(lw), O._p(truth), O._p(gw) if training else None, ns, training, O._p(cnt))
    wrong_want = O.lib.orc_wrong_count(n, C, O._p(lw), O._p(truth), None)
    d_l, d_t = to_dev(logits, dev), to_dev(truth, dev)
    d_g = torch.full((n, C), float("nan"), device=dev) if training else None
    res = torch.zeros(4, device=dev)
    ws = gcnb.zeroed_workspace(gcnb.lib.gcnb_ce_workspace(n), dev)
    for _ in range(2):  # second launch checks the self-resetting ticket
        d_l.copy_(to_dev(logits, dev))
        gcnb.softmax_ce(d_l, d_g, d_t, n, C, ns, training, res, ws)
    r = to_np(res)
    assert_close(r[0], loss_want, rtol=1e-5, what="loss sum")
    assert int(r.view(u32)[1]) == wrong_want and int(r.view(u32)[2]) == int(cnt[0])
    assert_close(to_np(d_l), lw, rtol=1e-6, atol=1e-6, what="shifted logits")
    if training:
        assert_close(to_np(d_g), gw, rtol=1e-5, atol=1e-9, what="ce grad")


@pytest.mark.parametrize("tensor_cores", [False, True])
@pytest.mark.parametrize("n,K,C,training,with_grad", [(5000, 16, 41, 1, 0), (5000, 16, 41, 0, 0), (20011, 16, 41, 1, 1), (2708, 16, 64, 1, 0),
                                                     (1000, 8, 9, 1, 1), (333, 32, 40, 1, 0), (31, 16, 17, 1, 1), (1, 16, 41, 1, 0),
                                                     (4096, 16, 20, 0, 0), (70000, 16, 41, 1, 0), (900, 16, 7, 1, 1)])
def test_output_head_in_one_kernel(O, gcnb, dev, n, K, C, training, with_grad, tensor_cores):
    """csrc/head.cu against the oracle's module chain: Matmul forward, cross-entropy + counts, Matmul backward."""
    import torch
    rng = np.random.default_rng(n * 7 + K + C)
    y = rng.standard_normal((n, K)).astype(f32)
    w = (rng.standard_normal((K, C)) * 0.7).astype(f32)
    truth = rng.integers(0, C, n).astype(i32)
    truth[rng.random(n) < 0.35] = -1
    ns = int((truth >= 0).sum()) + 3
    z = np.empty((n, C), f32)
    O.lib.orc_matmul(n, K, C, O._p(y), O._p(w), O._p(z))
    gw = np.zeros((n, C), f32)
    cnt = np.zeros(1, np.int64)
    loss_want = O.lib.orc_cross_entropy(n, C, O._p(z), O._p(truth), O._p(gw) if training else None, ns, training, O._p(cnt))
    wrong_want = O.lib.orc_wrong_count(n, C, O._p(z), O._p(truth), None)
    dy_want, dw_want = np.empty((n, K), f32), np.empty((K, C), f32)
    if training:
        O.lib.orc_matmul_bwd(n, K, C, O._p(y), O._p(w), O._p(gw), O._p(dy_want), O._p(dw_want))
    assert gcnb.lib.gcnb_head_supported(K, C) == 1
    if tensor_cores and not gcnb.lib.gcnb_head_tc_supported(K, C):
        pytest.skip("the tensor-core variant covers in_dim 16, classes <= 48")
    d_y, d_w, d_t = to_dev(y, dev), to_dev(w, dev), to_dev(truth, dev)
    d_z = torch.full((n, C), float("nan"), device=dev)
    d_g = torch.full((n, C), float("nan"), device=dev) if with_grad else None
    d_dy = torch.full((n, K), float("nan"), device=dev)
    d_dw = torch.full((K, C), float("nan"), device=dev)
    res = torch.zeros(4, device=dev)
    ws = gcnb.zeroed_workspace(gcnb.lib.gcnb_head_workspace(n, K, C), dev)
    outs = []
    for _ in range(2):  # second launch: self-resetting ticket, bit-repeatable sums
        gcnb.head(d_y, d_w, d_t, n, K, C, ns, training, d_z, d_g, d_dy, d_dw, res, ws, tensor_cores=tensor_cores)
        torch.cuda.synchronize()
        outs.append((to_np(res).copy(), to_np(d_dw).copy()))
    r = outs[0][0]
    assert (outs[0][0].view(u32) == outs[1][0].view(u32)).all()
    assert_close(r[0], loss_want, rtol=1e-5, what="loss sum")
    assert int(r.view(u32)[1]) == wrong_want and int(r.view(u32)[2]) == int(cnt[0])
    assert_close(to_np(d_z), z, rtol=1e-5, atol=1e-5, what="shifted logits")
    assert (to_np(d_z).argmax(1) == z.argmax(1)).mean() > 0.999
    if training:
        assert (outs[0][1].view(u32) == outs[1][1].view(u32)).all(), "weight gradient must be deterministic"
        assert_close(to_np(d_dy), dy_want, what="head dy")
        assert_close(to_np(d_dw), dw_want, rtol=2e-5, atol=5e-6 * np.abs(dw_want).max(), what="head dW")  # (the oracle adds 70 k fp32 terms in sequence)
        if with_grad:
            assert_close(to_np(d_g), gw, what="head dz")
    # the unfused kernels compute the same elements with the same arithmetic
    z2 = torch.empty((n, C), device=dev)
    gcnb.matmul_nn(d_y, d_w, z2, n, K, C)
    g2 = torch.empty((n, C), device=dev)
    res2 = torch.zeros(4, device=dev)
    ws2 = gcnb.zeroed_workspace(gcnb.lib.gcnb_ce_workspace(n), dev)
    gcnb.softmax_ce(z2, g2, d_t, n, C, ns, training, res2, ws2)
    # head_tc_kernel: split-TF32 products (fp32-level accuracy, other bits); the FMA kernel's element arithmetic is that of the
    # unfused kernels
    if tensor_cores:
        assert_close(to_np(d_z), to_np(z2), rtol=2e-6, atol=2e-6 * float(np.abs(z).max()), what="fused vs unfused logits")
    else:
        assert torch.equal(z2, d_z), "fused and unfused logits differ"
    if training:
        dy2 = torch.empty((n, K), device=dev)
        gcnb.matmul_nt(g2, d_w, dy2, n, K, C)
        lab = torch.from_numpy(truth >= 0).to(dev)
        if tensor_cores:
            assert_close(to_np(d_dy[lab]), to_np(dy2[lab]), rtol=1e-5, what="fused vs unfused dy")
            assert float(d_dy[~lab].abs().max()) == 0.0 if bool((~lab).any()) else True
        else:
            assert torch.equal(dy2[lab], d_dy[lab]), "fused and unfused dy differ"
        if with_grad:
            if tensor_cores:
                assert_close(to_np(d_g), to_np(g2), rtol=1e-5, what="fused vs unfused dz")
            else:
                assert torch.equal(g2, d_g)


def test_adam_and_sumsq(O, gcnb, dev):
    import torch
    rng = np.random.default_rng(5)
    sizes, decays = [1433 * 16, 16 * 7, 5], [True, False, True]
    ws = [rng.standard_normal(s).astype(f32) * 0.1 for s in sizes]
    ms = [np.zeros(s, f32) for s in sizes]
    vs = [np.zeros(s, f32) for s in sizes]
    d = [[to_dev(a, dev) for a in (w, m, v)] for w, m, v in zip(ws, ms, vs)]
    for step in range(1, 4):
        gs = [rng.standard_normal(s).astype(f32) * 0.01 for s in sizes]
        ss = O.lib.orc_adam_step_size(0.01, 0.9, 0.999, step)
        for w, g, m, v, dec in zip(ws, gs, ms, vs, decays):
            O.lib.orc_adam_step(w.size, O._p(w), O._p(g), O._p(m), O._p(v), int(dec), 5e-4, 0.9, 0.999, 1e-8, ss)
        gcnb.adam_step([(dw, to_dev(g, dev), dm, dv, dec) for (dw, dm, dv), g, dec in zip(d, gs, decays)], 5e-4, 0.9, 0.999, 1e-8, ss)
    for (dw, dm, dv), w, m, v in zip(d, ws, ms, vs):
        assert_close(to_np(dw), w, rtol=1e-6, what="adam w")
        assert_close(to_np(dm), m, rtol=1e-6, what="adam m")
        assert_close(to_np(dv), v, rtol=1e-6, what="adam v")
    out = torch.zeros(1, device=dev)
    wsb = gcnb.zeroed_workspace(gcnb.lib.gcnb_sumsq_workspace(sizes[0]), dev)
    for _ in range(2):
        gcnb.sumsq(d[0][0], out, wsb)
    assert_close(to_np(out)[0], O.lib.orc_sumsq(sizes[0], O._p(ws[0])), rtol=1e-5, what="sumsq")


def test_set_truth(O, gcnb, dev, datasets):
    import torch
    ds = datasets["citeseer"]
    for cur in (1, 2, 3):
        want = np.empty(ds.num_nodes, i32)
        O.lib.orc_set_truth(ds.num_nodes, O._p(ds.split), O._p(ds.label), cur, O._p(want))
        t = torch.empty(ds.num_nodes, dtype=torch.int32, device=dev)
        gcnb.set_truth(t, to_dev(ds.split, dev), to_dev(ds.label, dev), cur)
        assert (to_np(t) == want).all()


@pytest.mark.parametrize("n,F,P,p,off", [(1003, 602, 16, 0.5, 0), (517, 24, 16, 0.5, 6), (300, 130, 32, 0.2, 0), (65, 640, 8, 0.9, 3),
                                         (9, 5, 16, 0.0, 0), (5000, 128, 16, 0.5, 2), (2049, 301, 16, 0.3, 1), (16, 8, 16, 0.5, 0)])
def test_dense_feature_products_with_bit_mask(O, gcnb, dev, n, F, P, p, off):
    """mask bits == Dropout's keep decisions (bit-exact); masked X*W and (masked X)^T*dH == oracle on the dropped copy."""
    import torch
    seed, hist = 19990304, [(F * P, 1), (n * F + off, 2)]
    rng = np.random.default_rng(n * F)
    X = rng.standard_normal((n, F)).astype(f32)
    W = rng.standard_normal((F, P)).astype(f32)
    dH = rng.standard_normal((n, P)).astype(f32)
    size = n * F
    full_mask = np.empty(off + size, u8)
    O.lib.orc_dropout_mask_philox(off + size, p, seed, O._p(_oracle_draws(hist, off + size)), 1, O._p(full_mask))
    mask = np.ascontiguousarray(full_mask[off:])
    Xd = X.copy().ravel()
    O.lib.orc_dropout_apply(size, O._p(Xd), O._p(mask), O.lib.orc_dropout_scale(p, 1))
    want_out, want_dW, dummy = np.empty((n, P), f32), np.empty((F, P), f32), np.empty((n, F), f32)
    O.lib.orc_matmul(n, F, P, O._p(Xd), O._p(W), O._p(want_out))
    O.lib.orc_matmul_bwd(n, F, P, O._p(Xd), O._p(W), O._p(dH), O._p(dummy), O._p(want_dW))
    assert gcnb.lib.gcnb_dense_feat_supported(F, P)
    d_X, d_W, d_dH = to_dev(X, dev), to_dev(W, dev), to_dev(dH, dev)
    words = gcnb.lib.gcnb_dropout_maskbits_words(n, F)
    bits = torch.full((words,), -1, dtype=torch.int32, device=dev)
    gcnb.dropout_maskbits(bits, n, F, p, gcnb.make_rng(seed, hist, elem_offset=off))
    # tile layout: 32 rows per tile, tile padded to a multiple of 4 words
    wpt = words // ((n + 31) // 32)
    tiles = np.unpackbits(to_np(bits).view(u8), bitorder="little").reshape(-1, wpt * 32)
    got_bits = tiles[:, : 32 * F].reshape(-1)[:size]
    assert (got_bits == mask).all()
    assert (tiles[:, 32 * F:] == 0).all() and (tiles[:, : 32 * F].reshape(-1)[size:] == 0).all()
    out = torch.full((n, P), float("nan"), device=dev)
    gcnb.dense_feat_fwd(d_X, bits, p, d_W, out, n, F, P)
    assert_close(to_np(out), want_out, what="dense_feat_fwd")
    dW = torch.full((F, P), float("nan"), device=dev)
    gcnb.dense_feat_tn(d_X, bits, p, d_dH, dW, n, F, P)
    assert_close(to_np(dW), want_dW, rtol=2e-5, what="dense_feat_tn")
    dW2 = torch.empty_like(dW)
    gcnb.dense_feat_tn(d_X, bits, p, d_dH, dW2, n, F, P)
    assert torch.equal(dW, dW2)
    # eval form: no mask
    O.lib.orc_matmul(n, F, P, O._p(X), O._p(W), O._p(want_out))
    gcnb.dense_feat_fwd(d_X, None, 0.0, d_W, out, n, F, P)
    assert_close(to_np(out), want_out, what="dense_feat_fwd eval")
