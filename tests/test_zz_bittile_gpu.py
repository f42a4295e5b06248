"""GPU parity of the bit-tile GraphSum (csrc/spmm_bittile.cu: tcgen05.mma on 128 x 64 bit-map tiles + remainder CSR)
against the oracle's CSR product, through the C ABI.  Tolerance: 1e-5 relative with the 1e-6 x max|want| floor (the
result differs from the CSR product by the rounding of s_i * s_j against 1/sqrtf(deg_i * deg_j) and by summation order).

The file sorts last on purpose: these are the newest kernels (every test here ran green on B200 in round 2,
profiles/r2a_checklist_log.txt)."""
import os

import numpy as np
import pytest

from tests.test_bittile_cpu import gcn_graph
from tests.util import assert_close, to_dev, to_np

# a wedged mbarrier pipeline must end the run, not hang it: pytest-timeout's thread method exits the process
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]
f32 = np.float32
# A shape is (columns per tile, 128-row blocks per item); (0, 0) = the default (items of 256 rows x 64 columns)
SHAPES = [(0, 0), (64, 1), (128, 1)]


def _check(O, gcnb, dev, indptr, indices, values, rs=None, cs=None, min_tile_nnz=0, seed=0, shape=(0, 0), want_ell=None):
    import torch
    n = len(indptr) - 1
    x = np.random.default_rng(seed).standard_normal((n, 16)).astype(f32)
    want = np.empty((n, 16), f32)
    O.lib.orc_spmm(n, 16, O._p(indptr), O._p(indices), O._p(values), O._p(x), O._p(want))
    plan = gcnb.BitTilePlan(indptr, indices, values, n, rs, cs, min_tile_nnz=min_tile_nnz, chunk_cols=shape[0],
                            row_blocks=shape[1])
    info = plan.info()
    assert info["tile_nnz"] + info["rem_nnz"] == len(indices)
    if want_ell is not None:
        assert bool(info["ell"]) == want_ell, info
    d_x = to_dev(x, dev)
    out = torch.full((n, 16), float("nan"), device=dev)
    plan.spmm16(d_x, out)
    torch.cuda.synchronize()
    assert_close(to_np(out), want, what="bit-tile GraphSum")
    out2 = torch.full((n, 16), float("nan"), device=dev)
    for _ in range(3):  # fixed summation order: bit-identical launch to launch (also exercises barrier phase re-use)
        plan.spmm16(d_x, out2)
    assert torch.equal(out, out2)
    plan.close()
    return info


@pytest.mark.parametrize("cfg", [dict(n=3000, comm=6, intra=40, inter=3, thr=64), dict(n=777, comm=2, intra=60, inter=2, thr=32),
                                 dict(n=20000, comm=5, intra=120, inter=10, thr=0)])
@pytest.mark.parametrize("shape", SHAPES)
def test_community_graph_matches_oracle(O, gcnb, dev, cfg, shape):
    rng = np.random.default_rng(cfg["n"])
    indptr, indices, values = gcn_graph(rng, cfg["n"], cfg["comm"], cfg["intra"], cfg["inter"])
    info = _check(O, gcnb, dev, indptr, indices, values, min_tile_nnz=cfg["thr"] * (1 if shape == (64, 1) else 2), seed=1,
                  shape=shape, want_ell=True)  # GraphSum values factor: the remainder is the pattern-only gather
    assert info["n_tiles"] > 0 and info["tile_nnz"] > 0.4 * len(indices)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("n,density", [(128, 0.3), (1000, 0.2), (1024, 0.05)])
def test_dense_pattern_with_explicit_scales(O, gcnb, dev, n, density, shape):
    """every entry in a tile (explicit row / column scales, threshold 1): up to 16 tiles per row block => all four
    accumulators, accumulate flag, A / B stage re-use; n = 1000 leaves a ragged last block and chunk"""
    rng = np.random.default_rng(n)
    M = rng.random((n, n)) < density
    rs = (0.5 + rng.random(n)).astype(f32)
    cs = (0.5 + rng.random(n)).astype(f32)
    rows, cols = np.nonzero(M)
    indptr = np.zeros(n + 1, np.uint32)
    indptr[1:] = np.cumsum(M.sum(1))
    values = (rs[rows] * cs[cols]).astype(f32)
    info = _check(O, gcnb, dev, indptr, cols.astype(np.uint32), values, rs, cs, min_tile_nnz=1, seed=2, shape=shape)
    assert info["rem_nnz"] == 0


def test_duplicates_missing_diagonals_unfactored_values_and_sparse_graphs(O, gcnb, dev):
    rng = np.random.default_rng(11)
    indptr, indices, values = gcn_graph(rng, 777, 3, 30, 2, dup=40, drop_diag=(5, 300, 776))
    _check(O, gcnb, dev, indptr, indices, values, min_tile_nnz=64, seed=3)
    values2 = values.copy()
    values2[::7] *= 1.5  # not s_i * s_j: those entries must keep their value (remainder)
    _check(O, gcnb, dev, indptr, indices, values2, min_tile_nnz=64, seed=4, want_ell=False)  # valued remainder: generic kernel
    # nothing dense enough: everything is remainder, no MMA launch
    n = 4000
    ip = np.arange(0, 3 * n + 1, 3, dtype=np.uint32)
    ix = rng.integers(0, n, 3 * n).astype(np.uint32)
    info = _check(O, gcnb, dev, ip, ix, rng.standard_normal(3 * n).astype(f32), seed=5)
    assert info["n_tiles"] == 0


def _same_plans(host, devp, what):
    a, b = host.arrays(), devp.arrays()
    for k in a:
        if isinstance(a[k], int):
            assert a[k] == b[k], (what, k, a[k], b[k])
        else:
            assert a[k].shape == b[k].shape, (what, k, a[k].shape, b[k].shape)
            bad = np.nonzero(a[k] != b[k])[0]
            assert bad.size == 0, (what, k, int(bad.size), int(bad[0]), a[k][bad[:4]], b[k][bad[:4]])


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("cfg", [dict(n=3000, comm=6, intra=40, inter=3, thr=64, dup=0), dict(n=777, comm=2, intra=60, inter=2, thr=32, dup=60),
                                 dict(n=20000, comm=5, intra=120, inter=10, thr=0, dup=500), dict(n=5000, comm=2, intra=300, inter=400, thr=1500, dup=0)])
def test_device_built_plan_is_the_host_built_plan(O, gcnb, dev, cfg, shape):
    """csrc/spmm_bittile_build.cu: the plan built on the GPU from the device CSR (bit maps by atomicOr, remainder compacted in
    row order, ELL bundles filled on the device) equals the host builder's array by array -- duplicate entries (the first
    occurrence owns the bit), ragged last blocks, rows above 256 remainder entries (wide ELL bundles)"""
    import torch
    rng = np.random.default_rng(cfg["n"] + 5)
    indptr, indices, values = gcn_graph(rng, cfg["n"], cfg["comm"], cfg["intra"], cfg["inter"], dup=cfg["dup"])
    n = cfg["n"]
    thr = cfg["thr"] * (1 if shape == (64, 1) else 2)
    host = gcnb.BitTilePlan(indptr, indices, values, n, min_tile_nnz=thr, chunk_cols=shape[0], row_blocks=shape[1])
    d_ip, d_ix, d_v = to_dev(indptr, dev), to_dev(indices, dev), to_dev(values, dev)
    devp = gcnb.BitTilePlan.from_device(d_ip, d_ix, d_v, n, n, min_tile_nnz=thr, chunk_cols=shape[0], row_blocks=shape[1])
    assert devp is not None, "a GraphSum matrix factors: the device builder must take it"
    assert host.info() == devp.info()
    assert host.info()["n_tiles"] > 0 and host.info()["ell"] == 1
    _same_plans(host, devp, cfg)
    x = to_dev(rng.standard_normal((n, 16)).astype(f32), dev)
    out_h = torch.full((n, 16), float("nan"), device=dev)
    out_d = torch.full((n, 16), float("nan"), device=dev)
    host.spmm16(x, out_h)
    devp.spmm16(x, out_d)
    assert torch.equal(out_h, out_d)
    # explicit scales + a rectangular row block of the same matrix (what a rank of the partitioned engine builds)
    a = host.arrays()
    s = a["col_scale"].view(f32)
    r0, r1 = n // 3, n // 3 + min(n // 2, 2000)
    ip_blk = (indptr[r0:r1 + 1] - indptr[r0]).astype(np.uint32)
    ix_blk, v_blk = indices[indptr[r0]:indptr[r1]], values[indptr[r0]:indptr[r1]]
    host_blk = gcnb.BitTilePlan(ip_blk, ix_blk, v_blk, n, s[r0:r1].copy(), s, min_tile_nnz=thr, chunk_cols=shape[0], row_blocks=shape[1])
    if host_blk.info()["n_tiles"] > 0:
        dev_blk = gcnb.BitTilePlan.from_device(to_dev(ip_blk, dev), to_dev(ix_blk, dev), to_dev(v_blk, dev), r1 - r0, n,
                                               to_dev(s[r0:r1].copy(), dev), to_dev(s, dev), min_tile_nnz=thr, chunk_cols=shape[0],
                                               row_blocks=shape[1])
        assert dev_blk is not None
        _same_plans(host_blk, dev_blk, (cfg, "row block"))
        dev_blk.close()
    host_blk.close()
    host.close()
    devp.close()


def test_device_built_plan_with_hub_rows_and_explicit_scales(gcnb, dev):
    """rows of 9000 and 20000 remainder entries (ELL parts with partial slots) and a pattern given with explicit scales"""
    rng = np.random.default_rng(17)
    n = 12000
    rows = []
    for i in range(n):
        if i in (7, 5000):
            cols = rng.choice(n, 9000 if i == 7 else 11999, replace=False)
        else:
            b0 = (i // 512) * 512
            cols = np.concatenate([b0 + rng.choice(min(512, n - b0), 40, replace=False), rng.integers(0, n, 3)])
        rows.append(cols.astype(np.uint32))
    indptr = np.zeros(n + 1, np.uint32)
    indptr[1:] = np.cumsum([len(r) for r in rows])
    indices = np.concatenate(rows)
    rs, cs = (0.5 + rng.random(n)).astype(f32), (0.5 + rng.random(n)).astype(f32)
    values = rs[np.repeat(np.arange(n), np.diff(indptr.astype(np.int64)))] * cs[indices]
    host = gcnb.BitTilePlan(indptr, indices, values, n, rs, cs)
    assert host.info()["ell"] == 1 and host.arrays()["ell_slots"] > 0
    for d_v in (to_dev(values, dev), None):  # valued, and the pattern alone
        devp = gcnb.BitTilePlan.from_device(to_dev(indptr, dev), to_dev(indices, dev), d_v, n, n, to_dev(rs, dev), to_dev(cs, dev))
        assert devp is not None
        _same_plans(host, devp, "hub rows")
        devp.close()
    host.close()


def test_device_builder_with_saturating_16_bit_counters(gcnb, dev):
    """column ranges whose 32-bit chunk histogram would not fit shared memory (> 3.2 M columns) count in saturating 16-bit
    halves; forced here on small graphs (GCNB_BTB_COUNTER16=1), odd and even chunk counts, a chunk that saturates"""
    rng = np.random.default_rng(23)
    cases = [gcn_graph(rng, 3000, 6, 40, 3, dup=30) + (64,), gcn_graph(rng, 4100, 3, 200, 10) + (0,)]
    # one row block x one chunk with > 0x7fff entries: 256 rows x 64 columns full of duplicates
    n = 600
    rows = np.repeat(np.arange(256), 64 * 3)
    cols = np.tile(np.tile(np.arange(64), 3), 256)
    extra_r, extra_c = np.arange(n), np.arange(n)  # a diagonal, so every row and column has a scale
    rr, cc = np.concatenate([extra_r, rows]), np.concatenate([extra_c, cols])
    order = np.argsort(rr, kind="stable")
    rr, cc = rr[order], cc[order]
    ip = np.zeros(n + 1, np.uint32)
    ip[1:] = np.cumsum(np.bincount(rr, minlength=n))
    sc = (0.5 + rng.random(n)).astype(f32)
    os.environ["GCNB_BTB_COUNTER16"] = "1"
    try:
        for indptr, indices, values, thr in cases:
            nn = len(indptr) - 1
            host = gcnb.BitTilePlan(indptr, indices, values, nn, min_tile_nnz=thr)
            devp = gcnb.BitTilePlan.from_device(to_dev(indptr, dev), to_dev(indices, dev), to_dev(values, dev), nn, nn, min_tile_nnz=thr)
            assert devp is not None
            _same_plans(host, devp, "16-bit counters")
            host.close()
            devp.close()
        host = gcnb.BitTilePlan(ip, cc.astype(np.uint32), None, n, sc, sc)
        assert host.info()["n_tiles"] >= 1 and host.info()["ell"] == 1
        devp = gcnb.BitTilePlan.from_device(to_dev(ip, dev), to_dev(cc.astype(np.uint32), dev), None, n, n, to_dev(sc, dev), to_dev(sc, dev))
        assert devp is not None
        _same_plans(host, devp, "saturated chunk")
        host.close()
        devp.close()
    finally:
        os.environ.pop("GCNB_BTB_COUNTER16", None)


def test_device_builder_hands_unfactored_matrices_to_the_host_builder(gcnb, dev):
    rng = np.random.default_rng(3)
    indptr, indices, values = gcn_graph(rng, 777, 3, 30, 2, dup=10, drop_diag=(5, 300))  # rows without a diagonal: no scale
    assert gcnb.BitTilePlan.from_device(to_dev(indptr, dev), to_dev(indices, dev), to_dev(values, dev), 777, 777, min_tile_nnz=64) is None
    indptr, indices, values = gcn_graph(rng, 777, 3, 30, 2)
    values[::7] *= 1.5
    assert gcnb.BitTilePlan.from_device(to_dev(indptr, dev), to_dev(indices, dev), to_dev(values, dev), 777, 777, min_tile_nnz=64) is None


@pytest.mark.parametrize("case", ["short", "mixed", "long"])
def test_pattern_only_ell_gather_matches_float64(gcnb, dev, case):
    """csrc/spmm_ell.cu on its own: R = diag(row_scale) * pattern * B2 (the remainder kernel of the bit-tile plans)"""
    import torch
    rng = np.random.default_rng({"short": 1, "mixed": 2, "long": 3}[case])
    if case == "short":
        n_rows, n_cols = 5003, 4000
        lens = rng.integers(0, 60, n_rows)
    elif case == "mixed":
        n_rows, n_cols = 3001, 50000
        lens = rng.integers(0, 700, n_rows)  # rows above 256 entries: wide bundles
        lens[::97] = 0
    else:
        n_rows, n_cols = 40, 3000
        lens = rng.integers(0, 30, n_rows)
        lens[[3, 17, 30]] = [20000, 8193, 9000]  # cut rows: parts, partial slots, combine kernel
    indptr = np.zeros(n_rows + 1, np.uint32)
    indptr[1:] = np.cumsum(lens)
    indices = rng.integers(0, n_cols, int(indptr[-1])).astype(np.uint32)
    B2 = np.zeros((n_cols + 1, 16), f32)
    B2[:n_cols] = rng.standard_normal((n_cols, 16)).astype(f32)
    rs = (0.5 + rng.random(n_rows)).astype(f32)
    want = np.zeros((n_rows, 16), np.float64)
    np.add.at(want, np.repeat(np.arange(n_rows), lens), B2[indices].astype(np.float64))
    want *= rs[:, None].astype(np.float64)
    plan = gcnb.EllPlan(indptr, indices, n_cols)
    d_B2, d_rs = to_dev(B2, dev), to_dev(rs, dev)
    out = torch.full((n_rows, 16), float("nan"), device=dev)
    plan.gather16(d_B2, d_rs, out)
    torch.cuda.synchronize()
    assert_close(to_np(out), want, what="ELL gather " + case)
    out2 = torch.full((n_rows, 16), float("nan"), device=dev)
    for _ in range(3):  # the ticket counter re-arms itself; fixed summation order => identical bits
        plan.gather16(d_B2, d_rs, out2)
    assert torch.equal(out, out2)
    plan.close()


def test_attached_plan_routes_only_matching_calls(O, gcnb, dev):
    import torch
    rng = np.random.default_rng(5)
    n = 3000
    indptr, indices, values = gcn_graph(rng, n, 6, 40, 3)
    d_ip, d_ix, d_v = (to_dev(a, dev) for a in (indptr, indices, values))
    plan = gcnb.SpmmPlan(d_ip, d_ix, n)
    x16, x41, x7 = torch.randn(n, 16, device=dev), torch.randn(n, 41, device=dev), torch.randn(n, 7, device=dev)
    base16, base41, base7 = torch.empty(n, 16, device=dev), torch.empty(n, 41, device=dev), torch.empty(n, 7, device=dev)
    plan.spmm(d_v, x16, base16, 16)
    plan.spmm(d_v, x41, base41, 41)
    plan.spmm(d_v, x7, base7, 7)
    bt = gcnb.BitTilePlan(indptr, indices, values, n, min_tile_nnz=64)
    plan.attach_bittile(bt, d_v)
    out16, out41, out7 = torch.empty(n, 16, device=dev), torch.full((n, 41), float("nan"), device=dev), torch.empty(n, 7, device=dev)
    plan.spmm(d_v, x16, out16, 16)
    plan.spmm(d_v, x41, out41, 41)
    plan.spmm(d_v, x7, out7, 7)
    assert torch.equal(out7, base7), "widths below 16 stay on the generic kernel"
    assert not torch.equal(out41, base41), "wider operands run as 16-column slabs through the bit tiles (41 = 3 slabs, the last shifted)"
    assert_close(to_np(out41), to_np(base41), what="width 41 through bit-tile slabs")
    assert not torch.equal(out16, base16), "width 16 went through the bit tiles (different rounding)"
    assert_close(to_np(out16), to_np(base16), what="attached bit-tile plan")
    other = d_v.clone()
    plan.spmm(other, x16, out16, 16)
    assert torch.equal(out16, base16), "another value array stays on the generic kernel"
    plan.attach_bittile(None, None)
    plan.spmm(d_v, x16, out16, 16)
    assert torch.equal(out16, base16)
    plan.close()
    bt.close()


def test_engine_training_with_bit_tiles_matches_default_path(gcnb, dev):
    import importlib
    eng = importlib.import_module("parallel_gcn_b200.engine")

    def run(flag):
        if flag is not None:
            os.environ["GCNB_BITTILE"] = flag
        try:
            # 8 communities of 2500 nodes, ~160 neighbours inside: 6 % dense blocks; small enough for CUDA-graph replay,
            # so the captured epoch contains the bit-tile fork / join as well
            ds = eng.synth_dataset(20000, 20000 * 100, 32, 6, n_blocks=8, seed=11)
            g = eng.GCN(ds, hidden_dims=(16,), dropouts=(0.5, 0.5))
            assert g.graph_bittile() == (flag is None) and g.uses_cuda_graph()
            hist = [(g.train_epoch(), g.eval(2)) for _ in range(4)]
            w = [g.weight(l) for l in range(2)]
            launches = g.launches_per_epoch()
            g.close()
            return hist, w, launches
        finally:
            os.environ.pop("GCNB_BITTILE", None)

    (h0, w0, l0), (h1, w1, l1) = run("0"), run(None)  # the path itself is asserted inside run()
    for ep, ((t0, v0), (t1, v1)) in enumerate(zip(h0, h1)):
        assert abs(t0[0] - t1[0]) <= 2e-5 * (1 + ep) * abs(t0[0]) and abs(v0[0] - v1[0]) <= 2e-5 * (1 + ep) * abs(v0[0])
    for a, b in zip(w0, w1):
        assert_close(b, a, rtol=1e-4, atol=1e-6, what="weights after 4 epochs")


def test_wide_operands_run_as_16_column_slabs(O, gcnb, dev):
    import torch
    rng = np.random.default_rng(9)
    n, dim, ld = 3000, 70, 96  # 70 = 4 slabs + a shifted last one; operands are column slabs of wider matrices
    indptr, indices, values = gcn_graph(rng, n, 6, 40, 3)
    x = rng.standard_normal((n, ld)).astype(f32)
    want = np.empty((n, dim), f32)
    xs = np.ascontiguousarray(x[:, 8:8 + dim])
    O.lib.orc_spmm(n, dim, O._p(indptr), O._p(indices), O._p(values), O._p(xs), O._p(want))
    bt = gcnb.BitTilePlan(indptr, indices, values, n, min_tile_nnz=64)
    d_x = to_dev(x, dev)
    out = torch.full((n, ld), float("nan"), device=dev)
    bt.spmm_ld(d_x, ld, out, ld, dim, b_off=8, c_off=4)
    torch.cuda.synchronize()
    got = to_np(out)
    assert_close(got[:, 4:4 + dim], want, what="bit-tile slabs")
    assert np.isnan(got[:, :4]).all() and np.isnan(got[:, 4 + dim:]).all(), "columns outside the slab were written"
    bt.close()


# ---- background staging of the window-staged GraphSum (spmm_stage.cu, gcnb_spmm_plan_stage_async_*): opt-in until it has
# ---- been run on a GPU (GCNB_TEST_ASYNC_STAGE=1)
def test_background_staging_attaches_the_same_plan(gcnb, dev):
    import torch
    from tests.test_stage_cpu import community_csr
    rng = np.random.default_rng(21)
    n = 20000
    indptr, indices = community_csr(rng, n, 5, 100, 0.8, ((77, 2500),))
    values = rng.standard_normal(len(indices)).astype(f32)
    d_ip, d_ix, d_v = (to_dev(a, dev) for a in (indptr, indices, values))
    x = torch.randn(n, 16, device=dev)
    sync_plan = gcnb.SpmmPlan(d_ip, d_ix, n)
    generic = torch.empty(n, 16, device=dev)
    sync_plan.spmm(d_v, x, generic, 16)
    assert sync_plan.stage(d_v, 16)["staged"] == 1
    staged = torch.empty(n, 16, device=dev)
    sync_plan.spmm(d_v, x, staged, 16)
    plan = gcnb.SpmmPlan(d_ip, d_ix, n)
    job = plan.stage_async_begin(d_v, 16)
    assert job is not None
    out = torch.empty(n, 16, device=dev)
    for _ in range(5):  # the plan serves products on the generic kernel while the helper builds
        plan.spmm(d_v, x, out, 16)
        assert torch.equal(out, generic)
    torch.cuda.synchronize()
    assert plan.stage_async_finish(job)["staged"] == 1
    plan.spmm(d_v, x, out, 16)
    assert torch.equal(out, staged), "the background build must produce the plan the synchronous call produces"
    plan.close()
    sync_plan.close()


def test_engine_background_setup_switches_at_a_fixed_epoch(gcnb, dev):
    """graphs too large for CUDA-graph replay build their static GraphSum representation on a helper thread and attach it
    before a FIXED training epoch: bit tiles when the graph has dense blocks, window staging otherwise (or with
    GCNB_BITTILE=0); a synchronous build (GCNB_ASYNC_STAGE=0) ends in the same representation"""
    import importlib
    eng = importlib.import_module("parallel_gcn_b200.engine")
    # > 8 Mi entries so that CUDA-graph replay is off and the background build applies
    # communities of 2500 nodes, 2.6 % dense: ~420 entries per 128 x 128 cells (bit tiles) and ~80 entries of a row per
    # window of 3072 columns (window staging, when bit tiles are refused: GCNB_BT_MIN_COVERAGE=101 makes the builder find
    # "no dense blocks worth it", the fallback inside the same helper thread)
    ds = eng.synth_dataset(60000, 60000 * 80, 16, 6, n_blocks=24, sigma=1.0, seed=3)

    def run(ds, env):
        os.environ.update(env)
        try:
            g = eng.GCN(ds, hidden_dims=(16,), dropouts=(0.5, 0.5))
            p0 = g.path_info()
            hist = [(g.train_epoch(), g.eval(2)) for _ in range(5)]
            p1 = g.path_info()
            w = [g.weight(l) for l in range(2)]
            g.close()
            return hist, w, p0, p1
        finally:
            for k in env:
                os.environ.pop(k, None)

    def same_curve(a, b):
        for ep, ((t0, v0), (t1, v1)) in enumerate(zip(a[0], b[0])):
            assert abs(t0[0] - t1[0]) <= 2e-5 * (1 + ep) * abs(t0[0]) and abs(v0[0] - v1[0]) <= 2e-5 * (1 + ep) * abs(v0[0])
        for x, y in zip(a[1], b[1]):
            assert_close(y, x, rtol=1e-4, atol=1e-6, what="weights after 5 epochs")

    sw = {"GCNB_STAGE_SWITCH_EPOCH": "2"}
    curves = []
    # (bit tiles come from the device builder by default -- nothing pending, see the end of this test; GCNB_BT_DEVICE_BUILD=0
    # keeps the host builder on its helper thread)
    for extra, kind in (({"GCNB_BT_DEVICE_BUILD": "0"}, "graph_bittile"), ({"GCNB_BT_MIN_COVERAGE": "101"}, "graph_staged"),
                        ({"GCNB_BITTILE": "0"}, "graph_staged")):
        other = "graph_staged" if kind == "graph_bittile" else "graph_bittile"
        sync = run(ds, dict(extra, GCNB_ASYNC_STAGE="0"))
        assert sync[2][kind] and sync[3][kind] and not sync[2][other] and not sync[2]["setup_pending"], (extra, sync[2], sync[3])
        bg1, bg2 = run(ds, dict(extra, **sw)), run(ds, dict(extra, **sw))
        assert bg1[2]["setup_pending"] and not bg1[2][kind] and bg1[3][kind] and not bg1[3][other] and not bg1[3]["setup_pending"], \
            (extra, bg1[2], bg1[3])
        assert bg1[0] == bg2[0] and all(np.array_equal(x, y) for x, y in zip(bg1[1], bg2[1])), "fixed switch epoch => reproducible bits"
        same_curve(sync, bg1)
        curves.append(sync)
    same_curve(curves[0], curves[1])
    same_curve(curves[1], curves[2])
    # default: the plan is built on the device inside the constructor -- bit tiles from the first epoch, the host builder's bits
    dflt = run(ds, {})
    assert dflt[2]["graph_bittile"] and not dflt[2]["setup_pending"] and dflt[3]["graph_bittile"], (dflt[2], dflt[3])
    assert dflt[0] == curves[0][0] and all(np.array_equal(x, y) for x, y in zip(dflt[1], curves[0][1])), \
        "device-built and host-built plans must give the same bits"


# ---- exact-split tcgen05 GEMM for the wide first layer (csrc/dense_tc.cu)
@pytest.mark.parametrize("n,f,p", [(128, 16, 16), (300, 50, 16), (1000, 602, 600), (4096 + 77, 602, 41)])
def test_exact_split_gemm_matches_float64(gcnb, dev, n, f, p):
    import torch
    rng = np.random.default_rng(n + f + p)
    X = rng.standard_normal((n, f)).astype(f32)
    W = ((rng.random((f, p)) - 0.5) * 2 * np.sqrt(6.0 / (f + p))).astype(f32)
    want = X.astype(np.float64) @ W.astype(np.float64)
    d_X, d_W = to_dev(X, dev), to_dev(W, dev)
    img = gcnb.dense_tc_pack_x(d_X, n, f)
    out = torch.full((n, p), float("nan"), device=dev)
    gcnb.dense_tc_fwd(img, d_W, out, n, f, p)
    torch.cuda.synchronize()
    assert_close(to_np(out), want, what="exact-split GEMM %dx%dx%d" % (n, f, p))
    out2 = torch.empty_like(out)
    gcnb.dense_tc_fwd(img, d_W, out2, n, f, p)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("n,f,p", [(200, 16, 16), (5000, 50, 41), (30000, 602, 600)])
def test_exact_split_weight_gradient_matches_float64(gcnb, dev, n, f, p):
    import torch
    rng = np.random.default_rng(n + f + p)
    X = rng.standard_normal((n, f)).astype(f32)
    dH = (rng.standard_normal((n, p)) * 1e-3).astype(f32)
    want = X.astype(np.float64).T @ dH.astype(np.float64)
    d_X, d_dH = to_dev(X, dev), to_dev(dH, dev)
    img = gcnb.dense_tc_pack_xt(d_X, n, f)
    dW = torch.full((f, p), float("nan"), device=dev)
    gcnb.dense_tc_tn(img, d_dH, dW, n, f, p)
    torch.cuda.synchronize()
    assert_close(to_np(dW), want, what="exact-split X^T dH %dx%dx%d" % (n, f, p))
    dW2 = torch.empty_like(dW)
    gcnb.dense_tc_tn(img, d_dH, dW2, n, f, p)
    assert torch.equal(dW, dW2)


# ---- engine parity on random ragged SYMMETRIC datasets (duplicate edges, explicit self entries, isolated and unlabelled
# ---- nodes, rows without features): opt-in (GCNB_TEST_RAGGED_ENGINE=1) until it has been run once on a GPU
def test_engine_matches_oracle_on_random_ragged_symmetric_datasets(O, gcnb, dev, tmp_path):
    import importlib
    eng = importlib.import_module("parallel_gcn_b200.engine")
    from tests.test_engine_gpu import _run_pair
    for seed in range(12):
        rng = np.random.default_rng(500 + seed)
        n = int(rng.integers(8, 200))
        rows = [[] for _ in range(n)]
        for _ in range(int(rng.integers(0, 4 * n))):
            i, j = int(rng.integers(0, n)), int(rng.integers(0, n))
            m = 2 if rng.random() < 0.1 else 1            # duplicate edges, both directions
            for _ in range(m):
                rows[i].append(j)
                if i != j:
                    rows[j].append(i)                      # i == j: an explicit self entry (listed once)
        root = tmp_path / ("g%d" % seed)
        d = root / "data"
        d.mkdir(parents=True)
        (d / "t.graph").write_text("".join(" ".join(map(str, r)) + "\n" for r in rows))
        lines = []
        for i in range(n):
            if rng.random() < 0.1:
                lines.append("\n")                         # unlabelled node without features
                continue
            idx = sorted(set(int(x) for x in rng.integers(0, 20, int(rng.integers(0, 6)))))
            lines.append(" ".join([str(int(rng.integers(0, 5)))] + ["%d:%g" % (k, rng.normal()) for k in idx]) + "\n")
        (d / "t.svmlight").write_text("".join(lines))
        (d / "t.split").write_text("".join("%d\n" % int(rng.integers(1, 4)) for _ in range(n)))
        ds_e = eng.parse_dataset(str(root), "t")
        ds_o = O.parse_dataset(str(d / "t"))
        if ds_e is None or not all(((ds_o.split == k) & (ds_o.label >= 0)).sum() > 0 for k in (1, 2, 3)):
            continue
        og, g, hist = _run_pair(O, eng, ds_o, ds_e, 3)
        for ep, (to, te, vo, ve) in enumerate(hist):
            assert abs(te[0] - to[0]) <= 2e-5 * (1 + ep) * abs(to[0]) and abs(ve[0] - vo[0]) <= 2e-5 * (1 + ep) * abs(vo[0]), (seed, ep)
            assert abs(te[1] - to[1]) < 1e-6 and abs(ve[1] - vo[1]) < 1e-6, (seed, ep)
        for l in range(2):
            assert_close(g.weight(l), og.W[l], rtol=1e-4, atol=1e-6, what="weights, dataset %d" % seed)
        g.close()
