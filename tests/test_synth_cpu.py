"""Row-local symmetric generator of the scale-out workloads (host/src/synth.cpp, gcnb_synth_sym_rows): what the
multi-GPU bench relies on -- any row block equals the same rows of the whole graph, the union of the blocks is a
symmetric simple graph in the parser's row convention, and graph values follow the reference's arithmetic."""
import importlib

import numpy as np
import pytest


@pytest.fixture(scope="module")
def eng():
    import __graft_entry__ as ge
    ge.load_package()
    return importlib.import_module("parallel_gcn_b200.engine")


def test_sym_rows_is_symmetric_simple_and_block_consistent(eng):
    n, kw = 12000, dict(block_size=500, mean_intra=30, mean_inter=8, n_reflect=256, sigma=1.0, seed=7)
    ip, ix = eng.synth_sym_rows(n, 0, n, **kw)
    ip64 = ip.astype(np.int64)
    rows = np.repeat(np.arange(n), np.diff(ip64))
    # row i = [i, strictly ascending neighbours without i]
    assert np.array_equal(ix[ip64[:-1]], np.arange(n, dtype=np.uint32))
    inner = np.ones(len(ix), bool)
    inner[ip64[:-1]] = False
    d = np.diff(ix.astype(np.int64))
    same_row = rows[1:] == rows[:-1]
    assert np.all(d[same_row & inner[1:] & inner[:-1]] > 0)
    assert not np.any((ix == rows) & inner)
    # symmetric: the multiset of (i, j) equals the multiset of (j, i)
    key = rows * n + ix.astype(np.int64)
    keyT = ix.astype(np.int64) * n + rows
    assert np.array_equal(np.sort(key), np.sort(keyT))
    # communities and skew are there
    intra = np.mean((rows // 500) == (ix // 500))
    deg = np.diff(ip64)
    assert 0.7 < intra < 0.9 and deg.max() > 3 * deg.mean()
    # any block generated on its own equals the slice (what every rank of a partitioned job does)
    for r0, r1 in ((0, 3000), (3000, 9004), (9004, n)):
        bp, bx = eng.synth_sym_rows(n, r0, r1 - r0, **kw)
        assert np.array_equal(bx, ix[ip64[r0]:ip64[r1]])
        assert np.array_equal(bp.astype(np.int64) + ip64[r0], ip64[r0:r1 + 1])
    # another seed gives another graph
    ip2, ix2 = eng.synth_sym_rows(n, 0, n, **dict(kw, seed=8))
    assert len(ix2) != len(ix) or not np.array_equal(ix2, ix)


def test_graph_values_follow_the_reference_arithmetic(eng):
    n = 3000
    ip, ix = eng.synth_sym_rows(n, 0, n, block_size=300, mean_intra=20, mean_inter=5, n_reflect=64, sigma=0.8, seed=3)
    deg = np.diff(ip.astype(np.int64)).astype(np.uint32)
    r0, r1 = 1000, 2200
    bp, bx = eng.synth_sym_rows(n, r0, r1 - r0, block_size=300, mean_intra=20, mean_inter=5, n_reflect=64, sigma=0.8, seed=3)
    gv = eng.synth_graph_values(bp, bx, r0, deg)
    rows = np.repeat(np.arange(r0, r1), np.diff(bp.astype(np.int64)))
    # src/parser.cpp:164-181: 1. / sqrtf(deg_src * deg_dst) (unsigned product -> float -> sqrtf -> double divide -> float)
    prod = (deg[rows] * deg[bx]).astype(np.float32)
    want = (1.0 / np.sqrt(prod, dtype=np.float32).astype(np.float64)).astype(np.float32)
    assert np.array_equal(gv, want)


def test_uniform_features_are_offset_consistent(eng):
    fp, fi, fv = eng.synth_dense_features_uniform(100, 12, 5, 0)
    fp2, fi2, fv2 = eng.synth_dense_features_uniform(40, 12, 5, 60 * 12)
    assert np.array_equal(fv[60 * 12:], fv2) and np.array_equal(fi2[:12], np.arange(12, dtype=np.uint32))
    assert fp[-1] == 1200 and abs(float(fv.mean())) < 0.15 and 0.8 < float(fv.std()) < 1.2


def test_local_inter_community_edges_stay_inside_the_window(eng):
    """gcnb_synth_sym_rows_local: edges that leave a community reach at most inter_window rows -- the halo-exchange case"""
    n, win = 12000, 1500
    kw = dict(block_size=500, mean_intra=30, mean_inter=8, n_reflect=256, sigma=1.0, seed=7, inter_window=win)
    ip, ix = eng.synth_sym_rows(n, 0, n, **kw)
    ip64 = ip.astype(np.int64)
    rows = np.repeat(np.arange(n), np.diff(ip64))
    key, keyT = rows * n + ix.astype(np.int64), ix.astype(np.int64) * n + rows
    assert np.array_equal(np.sort(key), np.sort(keyT))                      # symmetric
    assert np.array_equal(ix[ip64[:-1]], np.arange(n, dtype=np.uint32))     # self entry first
    leaves = (rows // 500) != (ix // 500)
    assert 0.05 < leaves.mean() < 0.4 and np.abs(rows - ix.astype(np.int64))[leaves].max() <= win
    for r0, r1 in ((0, 4000), (4000, n)):                                    # row blocks generated on their own
        bp, bx = eng.synth_sym_rows(n, r0, r1 - r0, **kw)
        assert np.array_equal(bx, ix[ip64[r0]:ip64[r1]])
    # a row block references only the rows within the window of its borders
    blk = ix[ip64[4000]:ip64[8000]].astype(np.int64)
    assert blk.min() >= 4000 - win and blk.max() < 8000 + win
