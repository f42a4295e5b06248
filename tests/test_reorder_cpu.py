"""Locality reordering (host/src/reorder.cpp): label-propagation communities -> renumbering.  Checked on a planted
community graph whose node ids were shuffled: the renumbered graph is the same graph, and the window-staging builder
finds on it what it finds on the original community order."""
import importlib

import numpy as np
import pytest


@pytest.fixture(scope="module")
def mods():
    import __graft_entry__ as ge
    ge.load_package()
    return importlib.import_module("parallel_gcn_b200.engine"), importlib.import_module("parallel_gcn_b200.binding")


def permute_csr(eng, ip, ix, perm):
    n = len(ip) - 1
    op, ox = np.empty(n + 1, np.uint32), np.empty(len(ix), np.uint32)
    eng.check(eng.lib.gcnb_permute_csr(n, eng._p(ip), eng._p(ix), eng._p(perm), eng._p(op), eng._p(ox)))
    return op, ox


def edge_keys(ip, ix):
    n = len(ip) - 1
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(ip.astype(np.int64)))
    return rows * n + ix.astype(np.int64)


def test_reordering_recovers_shuffled_communities(mods):
    eng, gcnb = mods
    n = 60000
    ip, ix = eng.synth_graph(n, n * 30, 12, 0.8, 1.0, 2000, 5)
    shuffle = np.random.default_rng(0).permutation(n).astype(np.uint32)
    sp, sx = permute_csr(eng, ip, ix, shuffle)
    # the shuffled graph is the same graph under the renumbering; rows keep the parser convention
    assert np.array_equal(np.sort(edge_keys(sp, sx)), np.sort(shuffle[np.repeat(np.arange(n), np.diff(ip.astype(np.int64)))].astype(np.int64) * n + shuffle[ix]))
    assert np.array_equal(sx[sp[:-1].astype(np.int64)], np.arange(n, dtype=np.uint32))

    def staged_fraction(a, b):
        return gcnb.stage_host_build(a, b, n, 16, 0, 0, 0, 0, 148, 4)["staged_nnz"] / len(b)

    f_orig, f_shuf = staged_fraction(ip, ix), staged_fraction(sp, sx)
    assert f_orig > 0.5 and f_shuf < 0.2
    perm, n_comm = eng.reorder_communities(sp, sx)
    assert np.array_equal(np.sort(perm), np.arange(n, dtype=np.uint32))      # a permutation
    assert 12 <= n_comm < 2000
    rp, rx = permute_csr(eng, sp, sx, perm)
    assert staged_fraction(rp, rx) > 0.95 * f_orig
    # deterministic, and independent of how the work is split (the library picks the thread count itself)
    perm2, _ = eng.reorder_communities(sp, sx)
    assert np.array_equal(perm, perm2)


def test_permute_rows_roundtrip(mods):
    eng, _ = mods
    rng = np.random.default_rng(1)
    n = 1000
    perm = rng.permutation(n).astype(np.uint32)
    a = rng.standard_normal((n, 7)).astype(np.float32)
    b = eng.permute_rows(a, perm)
    assert np.array_equal(b[perm], a)
    assert np.array_equal(eng.permute_rows(b, perm, inverse=True), a)
    lab = rng.integers(0, 5, n).astype(np.int32)
    assert np.array_equal(eng.permute_rows(lab, perm)[perm], lab)
