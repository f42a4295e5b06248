"""Pattern-only ELL remainder plan (parallel-gcn_b200/csrc/spmm_ell.cu, host part) checked WITHOUT a GPU: a numpy
emulation walks the bundles exactly as ell_gather16_kernel does (uint4-rows of 8 lane groups x 4 indices, padding index =
the zero row, wide bundles reduced over the groups, partial slots of cut rows added in ascending order) and must
reproduce pattern * B2 scaled by the rows; every CSR entry must appear exactly once."""
import numpy as np
import pytest

from tests.util import assert_close


@pytest.fixture(scope="module")
def gcnb():
    import importlib
    import __graft_entry__ as ge
    ge.load_package()
    return importlib.import_module("parallel_gcn_b200.binding")


def emulate(plan, B2, row_scale):
    n_rows = plan["n_rows"]
    R = np.full((n_rows, 16), np.nan, np.float32)
    slots = np.full((max(1, plan["n_slots"]), 16), np.nan, np.float32)
    idx = plan["idx"].reshape(-1, 8, 4)  # [uint4-row][group][k % 4]
    seen = []
    for b in range(plan["n_bundles"]):
        s4 = int(plan["steps"][b] & 0x7fffffff)
        wide = bool(plan["steps"][b] >> 31)
        blk = idx[plan["off"][b]:plan["off"][b] + s4]                    # (s4, 8, 4)
        assert plan["off"][b + 1] - plan["off"][b] == s4
        per_group = blk.transpose(1, 0, 2).reshape(8, s4 * 4)            # group g: its index stream in step order
        acc = np.zeros((8, 16), np.float32)
        for k in range(s4 * 4):                                          # sequential fp32 adds, as the kernel
            acc += B2[per_group[:, k]]
        rows = plan["rows"][b * 8:b * 8 + 8]
        if wide:
            for o in (1, 2, 4):                                          # xor tree over the groups (lane offsets 4, 8, 16)
                acc = acc + acc[np.arange(8) ^ o]
            row, slot = int(rows[0]), int(rows[1])
            if slot == 0xffffffff:
                assert np.isnan(R[row]).all()
                R[row] = acc[0] * row_scale[row]
            else:
                slots[slot] = acc[0]
            seen.append((np.full(per_group.size, row), per_group.reshape(-1)))
        else:
            for g in range(8):
                if rows[g] != 0xffffffff:
                    assert np.isnan(R[rows[g]]).all()
                    R[rows[g]] = acc[g] * row_scale[rows[g]]
                    seen.append((np.full(s4 * 4, rows[g]), per_group[g]))
                else:
                    assert (per_group[g] == plan["n_cols"]).all()
    for k in range(plan["n_split"]):
        row = plan["split_row"][k]
        a = np.zeros(16, np.float32)
        for s in range(plan["split_ptr"][k], plan["split_ptr"][k + 1]):
            a += slots[s]
        assert np.isnan(R[row]).all()
        R[row] = a * row_scale[row]
    return R, seen


@pytest.mark.parametrize("case", ["short", "mixed", "long", "empty"])
def test_ell_plan_reproduces_the_pattern_product(gcnb, case):
    rng = np.random.default_rng(hash(case) % 1000)
    if case == "short":
        n_rows, n_cols = 203, 150
        lens = rng.integers(0, 40, n_rows)
    elif case == "mixed":
        n_rows, n_cols = 97, 3000
        lens = rng.integers(0, 600, n_rows)     # rows above 64 entries become wide bundles (plans with few rows)
        lens[5] = 0
    elif case == "long":
        n_rows, n_cols = 14, 500
        lens = np.array([20000, 3, 9000, 0, 257, 256, 8192, 8193, 1, 300, 17000, 5, 64, 65])  # cut rows: parts + slots
    else:
        n_rows, n_cols, lens = 9, 4, np.zeros(9, np.int64)
    indptr = np.zeros(n_rows + 1, np.uint32)
    indptr[1:] = np.cumsum(lens)
    indices = rng.integers(0, n_cols, int(indptr[-1])).astype(np.uint32)  # duplicates allowed
    plan = gcnb.ell_host_build(indptr, indices, n_cols)
    assert plan["nnz"] == indptr[-1] and plan["wide_min"] == 64  # few rows: the 8 lane groups share rows above 64 entries
    B2 = np.zeros((n_cols + 1, 16), np.float32)
    B2[:n_cols] = rng.standard_normal((n_cols, 16)).astype(np.float32)
    rs = (0.5 + rng.random(n_rows)).astype(np.float32)
    R, seen = emulate(plan, B2, rs)
    assert not np.isnan(R).any()                                          # every row is written exactly once
    # every CSR entry exactly once, everything else is padding
    if seen:
        rr, cc = np.concatenate([s[0] for s in seen]), np.concatenate([s[1] for s in seen])
        keep = cc != n_cols
        got = np.sort(rr[keep].astype(np.int64) * (n_cols + 1) + cc[keep])
    else:
        got = np.empty(0, np.int64)
    want = np.sort(np.repeat(np.arange(n_rows), lens).astype(np.int64) * (n_cols + 1) + indices)
    assert np.array_equal(got, want)
    ref = np.zeros((n_rows, 16), np.float64)
    np.add.at(ref, np.repeat(np.arange(n_rows), lens), B2[indices].astype(np.float64))
    assert_close(R, ref * rs[:, None].astype(np.float64), what="ELL emulation " + case)
    # ticket order = longest bundle first inside each kind
    st = plan["steps"]
    wide = st >> 31
    assert (np.diff(wide.astype(np.int64)) <= 0).all()
    for kind in (0, 1):
        s = (st[wide == kind] & 0x7fffffff).astype(np.int64)
        assert (np.diff(s) <= 0).all()


def test_ell_plans_with_many_rows_keep_eight_rows_per_warp(gcnb):
    """from 65536 rows on the narrow bundles alone fill the machine: only rows above 256 entries are shared by the lane groups"""
    rng = np.random.default_rng(8)
    n_rows, n_cols = 70000, 5000
    lens = rng.integers(0, 12, n_rows)
    lens[[3, 500, 69999]] = [256, 257, 100]
    indptr = np.zeros(n_rows + 1, np.uint32)
    indptr[1:] = np.cumsum(lens)
    indices = rng.integers(0, n_cols, int(indptr[-1])).astype(np.uint32)
    plan = gcnb.ell_host_build(indptr, indices, n_cols)
    assert plan["wide_min"] == 256
    wide = plan["steps"] >> 31
    assert int(wide.sum()) == 1 and plan["rows"][0] == 500  # the one row above 256 entries, first in ticket order


def test_bittile_plan_counts_unfactored_entries(gcnb):
    """the bit-tile builder reports whether every entry factors (only then may the remainder drop its values)"""
    from tests.test_bittile_cpu import gcn_graph
    rng = np.random.default_rng(3)
    indptr, indices, values = gcn_graph(rng, 900, 3, 60, 3)
    plan = gcnb.bittile_host_build(indptr, indices, values, 900, min_tile_nnz=64)
    assert plan["n_unfactored"] == 0
    values2 = values.copy()
    values2[7] *= 1.01
    values2[-1] = 0.123
    plan2 = gcnb.bittile_host_build(indptr, indices, values2, 900, min_tile_nnz=64)
    assert plan2["n_unfactored"] == 2
