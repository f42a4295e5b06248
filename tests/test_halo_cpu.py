"""Host logic of the multi-GPU data layout (no GPU): the balanced, community-aligned partitioner
(host/src/reorder.cpp: gcnb_partition_communities) and the send lists of the halo exchange
(csrc/comm.cu: gcnb_halo_lists_from_masks), checked on planted-community graphs whose node ids were shuffled."""
import ctypes as C
import importlib

import numpy as np
import pytest


@pytest.fixture(scope="module")
def mods():
    import __graft_entry__ as ge
    ge.load_package()
    return (importlib.import_module("parallel_gcn_b200.engine"), importlib.import_module("parallel_gcn_b200.binding"),
            importlib.import_module("parallel_gcn_b200.dist"))


def shuffled_community_graph(eng, n, deg, blocks, seed):
    ip, ix = eng.synth_graph(n, n * deg // 2, blocks, 0.9, 0.8, 800, seed)
    shuffle = np.random.default_rng(seed).permutation(n).astype(np.uint32)
    sp, sx = np.empty(n + 1, np.uint32), np.empty(len(ix), np.uint32)
    eng.check(eng.lib.gcnb_permute_csr(n, eng._p(ip), eng._p(ix), eng._p(shuffle), eng._p(sp), eng._p(sx)))
    return sp, sx, shuffle


@pytest.mark.parametrize("world", [2, 3, 8])
def test_partition_is_balanced_and_cuts_on_community_borders(mods, world):
    eng, _, _ = mods
    n, blocks = 40000, 24
    ip, ix, _ = shuffled_community_graph(eng, n, 40, blocks, 7)
    new_of_old, block, rows, st = eng.partition_communities(ip, ix, world)
    assert block % 4 == 0 and sum(rows) == n and max(rows) <= block and len(rows) == world
    assert len(np.unique(new_of_old)) == n and new_of_old.max() < world * block
    rank = new_of_old // block
    assert np.array_equal(np.bincount(rank, minlength=world), rows)
    assert ((new_of_old % block) < np.asarray(rows)[rank]).all()                 # ids inside the rank's used range
    # balanced in CSR entries (GraphSum work), within the snapping tolerance
    deg = np.diff(ip.astype(np.int64))
    per_rank = np.bincount(rank, weights=deg, minlength=world)
    assert per_rank.max() <= 1.12 * deg.sum() / world, per_rank
    assert st["max_rank_entries"] == int(per_rank.max()) and st["total_entries"] == int(deg.sum())
    # fewer cut entries than equal row blocks of the (shuffled) numbering it was given; close to what the planted
    # communities allow: 10 % of the edges leave their community by construction
    src = np.repeat(np.arange(n), deg)
    cut = int((rank[src] != rank[ix]).sum())
    assert cut == st["cut_entries"]
    eq = (n + world - 1) // world
    assert int(((src // eq) != (ix // eq)).sum()) == st["cut_entries_equal_row_blocks"]
    assert cut < 0.35 * st["cut_entries_equal_row_blocks"], st
    assert blocks <= st["communities"] < 2000
    # ... and close to what the planted structure allows: 10 % of the edges leave their community by construction, a share
    # (world - 1) / world of those has to cross ranks
    assert cut <= 1.25 * 0.10 * (world - 1) / world * st["total_entries"], (cut, st)
    # deterministic
    again = eng.partition_communities(ip, ix, world)[0]
    assert np.array_equal(again, new_of_old)


def test_balanced_partition_gives_every_rank_an_ordinary_row_block(mods):
    eng, _, dist = mods
    n, world = 9000, 4
    ip, ix, _ = shuffled_community_graph(eng, n, 24, 10, 3)
    rng = np.random.default_rng(1)
    F = 12
    ds = eng.HostDataset(g_indptr=ip, g_indices=ix, f_indptr=(np.arange(n + 1) * F).astype(np.uint32),
                         f_indices=np.tile(np.arange(F, dtype=np.uint32), n), f_value=rng.standard_normal(n * F).astype(np.float32),
                         label=rng.integers(0, 5, n).astype(np.int32), split=rng.integers(1, 4, n).astype(np.uint32),
                         input_dim=F, output_dim=5)
    out, new_of_old, info = eng.balanced_partition(ds, world)
    n_pad = world * info["block"]
    assert out.num_nodes == n_pad and sum(info["rows"]) == n
    # the same graph under the renumbering, plus isolated dummy nodes (a self entry, no label, no split, zero features)
    deg_new = np.diff(out.g_indptr.astype(np.int64))
    assert np.array_equal(deg_new[new_of_old], np.diff(ip.astype(np.int64)))
    dummy = np.ones(n_pad, bool)
    dummy[new_of_old] = False
    assert (deg_new[dummy] == 1).all() and (out.label[dummy] == -1).all() and (out.split[dummy] == 0).all()
    assert np.array_equal(out.g_indices[out.g_indptr[:-1].astype(np.int64)], np.arange(n_pad, dtype=np.uint32))  # self first
    assert np.array_equal(out.label[new_of_old], ds.label) and np.array_equal(out.split[new_of_old], ds.split)
    assert np.array_equal(out.f_value.reshape(n_pad, F)[new_of_old], ds.f_value.reshape(n, F))
    assert (out.f_value.reshape(n_pad, F)[dummy] == 0).all()
    src = np.repeat(np.arange(n), np.diff(ip.astype(np.int64)))
    want = np.sort(new_of_old[src].astype(np.int64) * n_pad + new_of_old[ix])
    src_new = np.repeat(np.arange(n_pad), deg_new)
    keep = ~dummy[src_new]
    assert np.array_equal(np.sort(src_new[keep] * n_pad + out.g_indices[keep]), want)
    # the engine's ordinary row blocks: whole used ranges, balanced entry counts
    parts = [dist.partition_dataset(out, r, world) for r in range(world)]
    assert all(p["block"] == info["block"] for p in parts)
    nnz = [len(p["g_indices"]) - (info["block"] - info["rows"][r]) for r, p in enumerate(parts)]
    assert max(nnz) <= 1.12 * sum(nnz) / world


@pytest.mark.parametrize("world", [2, 5])
def test_halo_send_lists_hold_exactly_the_referenced_rows(mods, world):
    eng, gcnb, dist = mods
    n = 7001
    ip, ix = eng.synth_graph(n, n * 6, 9, 0.95, 0.7, 300, 11)
    block = dist.block_rows(n, world)
    words = (world * block + 31) // 32
    masks = np.zeros((world, words), np.uint32)
    rows_of = []
    for r in range(world):
        r0, r1 = min(n, r * block), min(n, (r + 1) * block)
        cols = ix[ip[r0]:ip[r1]].astype(np.int64)                                 # slot layout == global ids for equal blocks
        np.bitwise_or.at(masks[r], cols >> 5, (np.uint32(1) << (cols & 31).astype(np.uint32)))
        rows_of.append((r0, r1, np.unique(cols)))
    fn = gcnb.lib.gcnb_halo_lists_from_masks
    fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    total_sent = 0
    for r in range(world):
        r0, r1, _ = rows_of[r]
        out, off = C.c_void_p(), (C.c_int64 * (world + 1))()
        assert fn(masks.ctypes.data, world, r, words, block, r1 - r0, C.byref(out), off) == 0
        lst = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint32)), shape=(max(1, off[world]),))[:off[world]].copy()
        eng.lib.gcnb_host_free(out)
        assert off[0] == 0 and off[r] == off[r + 1]                               # nothing is sent to oneself
        for p in range(world):
            got = lst[off[p]:off[p + 1]]
            if p == r:
                continue
            cols_p = rows_of[p][2]
            want = cols_p[(cols_p >= r0) & (cols_p < r1)] - r0                    # rows of r that p's block references
            assert np.array_equal(got, want.astype(np.uint32)), (r, p)
        total_sent += off[world]
    full = sum((world - 1) * (b - a) for a, b, _ in rows_of)
    assert 0 < total_sent < full                                                  # a graph with locality ships less than whole slabs
