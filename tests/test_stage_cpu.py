"""Window-staging plan builder (parallel-gcn_b200/csrc/spmm_stage.cu, host part) checked WITHOUT a GPU: a numpy emulation
consumes the packed arrays exactly the way spmm_staged16_kernel does (same chunk / slot / step addressing) and must
reproduce the CSR product; every CSR entry must appear exactly once in (staged segments + remainder CSR)."""
import numpy as np
import pytest

from tests.util import assert_close


@pytest.fixture(scope="module")
def gcnb():
    import importlib
    import __graft_entry__ as ge
    ge.load_package()
    return importlib.import_module("parallel_gcn_b200.binding")


def community_csr(rng, n, n_comm, mean_deg, intra, heavy=()):
    deg = np.maximum(1, rng.poisson(mean_deg, n)).astype(np.int64)
    for r, d in heavy:
        deg[r] = d
    indptr = np.zeros(n + 1, np.uint32)
    indptr[1:] = np.cumsum(deg)
    rows = np.repeat(np.arange(n), deg)
    bs = (n + n_comm - 1) // n_comm
    local = np.minimum((rows // bs) * bs + rng.integers(0, bs, rows.size), n - 1)
    glob = rng.integers(0, n, rows.size)
    cols = np.where(rng.random(rows.size) < intra, local, glob).astype(np.uint32)
    return indptr, cols


def emulate(plan, indptr, indices, values, B):
    """what the device does: lane l of a bundle walks its own segment (block / lane / step addressing of
    spmm_staged16_kernel), partial row per slot; remainder product; then the row's slots added in slot_list order."""
    n_rows, dim = plan["n_rows"], B.shape[1]
    wc = plan["window_rows"]
    bundles, runs, lens = plan["bundles"], plan["runs"], plan["lens"]
    partial = np.zeros((plan["n_slots"], dim), np.float64)
    slot_done = np.zeros(plan["n_slots"], np.int64)
    row_of_slot = np.repeat(np.arange(n_rows), np.diff(plan["row_slot"].astype(np.int64)))
    seen = np.zeros(indices.size, np.int64)
    bundle_seen = np.zeros(len(bundles), np.int64)
    assert plan["run_begin"][0] == 0 and plan["run_begin"][-1] == len(runs)
    assert np.all(np.diff(plan["run_begin"].astype(np.int64)) >= 0)
    for w, bb, be, _ in runs:
        assert bb < be
        for b in range(bb, be):
            bundle_seen[b] += 1
            blk0, L, minL, _z = bundles[b]
            ln = lens[b * 32:(b + 1) * 32].astype(np.int64)
            assert ln.max() == L and (ln.min() == minL)
            for l in range(32):
                k = np.arange(ln[l])
                pos = ((blk0 + k // 4).astype(np.int64) * 32 + l) * 4 + k % 4
                lcol = plan["pidx"][pos].astype(np.int64)
                pe = plan["pperm"][pos].astype(np.int64)
                slot = plan["lane_slot"][b * 32 + l]
                if ln[l] == 0:
                    assert slot == 0xFFFFFFFF
                    continue
                row = row_of_slot[slot]
                slot_done[slot] += 1
                assert np.all(pe != 0xFFFFFFFF) and np.all(lcol < wc)
                gcol = int(w) * wc + lcol
                assert np.array_equal(indices[pe], gcol), "packed column id does not match the CSR entry"
                assert np.all((pe >= indptr[row]) & (pe < indptr[row + 1])), "entry of another row"
                seen[pe] += 1
                partial[slot] = (values[pe, None].astype(np.float64) * B[gcol]).sum(0)
                # padding of this lane up to the bundle's block count must be marked
                nblk = (L + 3) // 4
                kk = np.arange(ln[l], nblk * 4)
                ppos = ((blk0 + kk // 4).astype(np.int64) * 32 + l) * 4 + kk % 4
                assert np.all(plan["pperm"][ppos] == 0xFFFFFFFF)
    assert np.all(bundle_seen == 1), "a bundle is not covered by exactly one run"
    out = np.zeros((n_rows, dim), np.float64)
    rp, ri, rperm = plan["r_indptr"], plan["r_indices"], plan["r_perm"]
    assert rp[-1] == len(ri) == plan["rem_nnz"]
    assert np.array_equal(indices[rperm], ri)
    seen[rperm] += 1
    rrows = np.repeat(np.arange(n_rows), np.diff(rp.astype(np.int64)))
    assert np.all((rperm >= indptr[rrows]) & (rperm < indptr[rrows + 1]))
    np.add.at(out, rrows, values[rperm, None].astype(np.float64) * B[ri])
    # every slot written exactly once; a row's slots are the contiguous range row_slot[r] .. row_slot[r+1]
    assert plan["n_slots"] == plan["n_segs"] == int((lens > 0).sum())
    assert np.all(slot_done == 1)
    np.add.at(out, row_of_slot, partial)
    assert np.all(seen == 1), "CSR entries covered %d..%d times" % (seen.min(), seen.max())
    return out


def conflict_fraction(plan):
    """share of (lane l, lane l+4) pairs of a quarter-warp that read rows of the SAME parity in a step (2-way bank
    conflict on their 16-byte chunk); only steps in which both lanes are active count."""
    bad = tot = 0
    lens = plan["lens"].astype(np.int64)
    for b, (blk0, L, minL, _z) in enumerate(plan["bundles"]):
        for q in range(4):
            for l in range(8 * q, 8 * q + 4):
                n = min(lens[b * 32 + l], lens[b * 32 + l + 4])
                if n == 0:
                    continue
                k = np.arange(n)
                p0 = plan["pidx"][((blk0 + k // 4).astype(np.int64) * 32 + l) * 4 + k % 4] & 1
                p1 = plan["pidx"][((blk0 + k // 4).astype(np.int64) * 32 + l + 4) * 4 + k % 4] & 1
                bad += int(np.sum(p0 == p1))
                tot += n
    return bad / max(tot, 1)


def reference(indptr, indices, values, B):
    n = len(indptr) - 1
    out = np.zeros((n, B.shape[1]), np.float64)
    rows = np.repeat(np.arange(n), np.diff(indptr.astype(np.int64)))
    np.add.at(out, rows, values[:, None].astype(np.float64) * B[indices])
    return out


@pytest.mark.parametrize("threads", [1, 3])
@pytest.mark.parametrize("cfg", [
    dict(n=3000, comm=6, deg=60, intra=0.8, window=512, min_seg=8, seg_cap=64, heavy=((5, 4000), (2999, 900))),
    dict(n=1000, comm=2, deg=30, intra=0.9, window=500, min_seg=4, seg_cap=128, heavy=((0, 0),)),
    dict(n=2048, comm=4, deg=40, intra=0.6, window=300, min_seg=6, seg_cap=512, heavy=()),
])
def test_stage_plan_covers_csr_and_reproduces_product(gcnb, cfg, threads):
    rng = np.random.default_rng(11)
    indptr, indices = community_csr(rng, cfg["n"], cfg["comm"], cfg["deg"], cfg["intra"],
                                    [h for h in cfg["heavy"] if h[1] > 0])
    values = rng.standard_normal(indices.size).astype(np.float32)
    B = rng.standard_normal((cfg["n"], 16)).astype(np.float32)
    plan = gcnb.stage_host_build(indptr, indices, cfg["n"], 16, cfg["window"], cfg["min_seg"], cfg["seg_cap"],
                                 min_window_nnz=1, n_cta=7, n_threads=threads)
    assert plan["staged_nnz"] + plan["rem_nnz"] == indices.size
    assert plan["staged_nnz"] > 0.3 * indices.size
    used = plan["lens"][plan["lens"] > 0]
    assert used.max() <= cfg["seg_cap"] and used.min() >= min(cfg["min_seg"], cfg["seg_cap"] // 2)
    got = emulate(plan, indptr, indices, values, B)
    assert_close(got, reference(indptr, indices, values, B), rtol=1e-9, what="staged emulation")
    assert conflict_fraction(plan) < (0.12 if cfg["deg"] >= 60 else 0.3)  # random parity leaves ~0.8/sqrt(len) unpaired


def test_stage_plan_is_independent_of_thread_count(gcnb):
    rng = np.random.default_rng(5)
    indptr, indices = community_csr(rng, 5000, 5, 50, 0.8)
    a = gcnb.stage_host_build(indptr, indices, 5000, 16, 1024, 8, 128, 1, 11, 1)
    b = gcnb.stage_host_build(indptr, indices, 5000, 16, 1024, 8, 128, 1, 11, 4)
    for k in ("bundles", "runs", "run_begin", "pidx", "pperm", "row_slot", "r_indptr", "r_indices", "r_perm", "lens",
              "lane_slot"):
        assert np.array_equal(a[k], b[k]), k


def test_stage_plan_without_locality_stages_nothing(gcnb):
    rng = np.random.default_rng(3)
    n = 20000
    indptr, indices = community_csr(rng, n, 1, 20, 0.0)
    plan = gcnb.stage_host_build(indptr, indices, n, 16, 256, 16, 512, 0, 8, 2)
    assert plan["staged_nnz"] == 0 and plan["n_segs"] == 0 and plan["n_runs"] == 0 and plan["n_bundles"] == 0
    assert np.array_equal(plan["r_indices"], indices) and np.array_equal(plan["r_indptr"], indptr)


def test_stage_plan_queues_are_balanced(gcnb):
    rng = np.random.default_rng(9)
    indptr, indices = community_csr(rng, 20000, 10, 80, 0.85)
    plan = gcnb.stage_host_build(indptr, indices, 20000, 16, 1024, 8, 256, 1, 16, 2)
    cost = np.zeros(16)
    for q in range(16):
        for w, sb, se, _ in plan["runs"][plan["run_begin"][q]:plan["run_begin"][q + 1]]:
            # the builder's cycle model: window copy + per-bundle overhead + steps
            cost[q] += 6000 + 80 * (se - sb) + 18.0 * plan["bundles"][sb:se, 1].astype(np.int64).sum()
    assert cost.min() > 0 and cost.max() / cost.mean() < 1.15, cost


def test_stage_plan_own_column_range_splits_runs(gcnb):
    """row-partitioned GraphSum: windows entirely inside the rank's own column range go to a second run list; the two
    lists together cover every bundle exactly once and reproduce the product."""
    rng = np.random.default_rng(13)
    n = 6000
    indptr, indices = community_csr(rng, n, 6, 50, 0.8)
    values = rng.standard_normal(indices.size).astype(np.float32)
    B = rng.standard_normal((n, 16)).astype(np.float32)
    own = (1900, 4100)
    plan = gcnb.stage_host_build(indptr, indices, n, 16, 512, 8, 64, 1, 7, 2, own_cols=own)
    assert plan["n_own_runs"] > 0 and plan["n_runs"] > 0
    for w, sb, se, _ in plan["own_runs"]:
        assert w * 512 >= own[0] and min(n, (w + 1) * 512) <= own[1] and sb < se
    for w, sb, se, _ in plan["runs"]:
        assert not (w * 512 >= own[0] and min(n, (w + 1) * 512) <= own[1])
    assert plan["own_run_begin"][0] == 0 and plan["own_run_begin"][-1] == plan["n_own_runs"]
    base = gcnb.stage_host_build(indptr, indices, n, 16, 512, 8, 64, 1, 7, 2)
    assert base["n_own_runs"] == 0 and np.array_equal(base["bundles"], plan["bundles"])
    merged = dict(plan)
    merged["runs"] = np.concatenate([plan["runs"], plan["own_runs"]])
    merged["run_begin"] = np.array([0] * 7 + [len(merged["runs"])], np.uint32)
    got = emulate(merged, indptr, indices, values, B)
    assert_close(got, reference(indptr, indices, values, B), rtol=1e-9, what="own + remote runs")
