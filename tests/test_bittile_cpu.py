"""Bit-tile GraphSum plan (parallel-gcn_b200/csrc/spmm_bittile.cu, host part) checked WITHOUT a GPU.

A numpy emulation consumes the plan arrays the way the kernels do -- the expanders' bit -> bf16-pair formula, the pack
kernel's byte layout read back through the shared-memory descriptor's addressing (K-major core matrices, LBO 768 /
SBO 128), the four rotating accumulators per row block, the (lo + mid) + hi piece order, the remainder CSR -- and must
reproduce the CSR product; every CSR entry must appear exactly once in (bit maps + remainder)."""
import numpy as np
import pytest

from tests.util import assert_close

ROWS, CHUNK, NACC = 128, 64, 4
KSTEP_BYTES, CHUNK_BYTES = 1536, 6144


@pytest.fixture(scope="module")
def gcnb():
    import importlib
    import __graft_entry__ as ge
    ge.load_package()
    return importlib.import_module("parallel_gcn_b200.binding")


def gcn_graph(rng, n, n_comm, deg_intra, deg_inter, dup=0, drop_diag=()):
    """symmetric community graph in the reference's row layout (self first, parser.cpp:59-112) with GraphSum values
    1/sqrtf(deg_i * deg_j) computed as parser.cpp:164-181 does; `dup` extra duplicate entries; rows in drop_diag lose
    their self entry."""
    bs = (n + n_comm - 1) // n_comm
    src = np.repeat(np.arange(n), deg_intra)
    dst = np.minimum((src // bs) * bs + rng.integers(0, bs, src.size), n - 1)
    s2 = np.repeat(np.arange(n), deg_inter)
    d2 = rng.integers(0, n, s2.size)
    a = np.concatenate([src, s2, dst, d2])
    b = np.concatenate([dst, d2, src, s2])
    keep = a != b
    key = np.unique(a[keep].astype(np.int64) * n + b[keep])          # sorted (row, neighbour) pairs, no duplicates
    key = np.concatenate([key, np.arange(n, dtype=np.int64) * n + np.arange(n)])
    pi, pj = key // n, key % n
    order = np.lexsort((np.where(pi == pj, -1, pj), pi))                # per row: self first, then ascending neighbours
    pi, pj = pi[order], pj[order]
    alive = np.ones(pi.size, bool)
    alive[np.nonzero((pi == pj) & np.isin(pi, np.asarray(drop_diag, np.int64)))[0]] = False
    pi, pj = pi[alive], pj[alive]
    if dup:
        pick = rng.integers(0, pi.size, dup)                             # duplicate entries, appended to their rows
        pi, pj = np.concatenate([pi, pi[pick]]), np.concatenate([pj, pj[pick]])
        order = np.argsort(pi, kind="stable")
        pi, pj = pi[order], pj[order]
    indptr = np.zeros(n + 1, np.uint32)
    indptr[1:] = np.cumsum(np.bincount(pi, minlength=n))
    indices = pj.astype(np.uint32)
    deg = np.diff(indptr.astype(np.int64))
    rr = np.repeat(np.arange(n), deg)
    prod = (deg[rr] * deg[indices]).astype(np.uint32).astype(np.float32)
    values = (1.0 / np.sqrt(prod).astype(np.float64)).astype(np.float32)
    return indptr, indices, values


def bf16_split(x):
    """truncating split of fp32 into three bf16 pieces (as uint16 bit patterns) -- bt_pack_kernel"""
    x = x.astype(np.float32)
    xb = x.view(np.uint32)
    hb = xb & np.uint32(0xFFFF0000)
    r1 = x - hb.view(np.float32)
    mb = r1.view(np.uint32) & np.uint32(0xFFFF0000)
    r2 = r1 - mb.view(np.float32)
    lb = r2.view(np.uint32) & np.uint32(0xFFFF0000)
    assert np.array_equal(hb.view(np.float32).astype(np.float64) + mb.view(np.float32) + lb.view(np.float32),
                          x.astype(np.float64)), "the three pieces must add up exactly"
    return [(p >> np.uint32(16)).astype(np.uint16) for p in (hb, mb, lb)]


def pack_b(B, col_scale):
    """bt_pack_kernel: thread (group of 8 rows of B, column) writes three 16-byte vectors into the chunk image"""
    n_cols = B.shape[0]
    n_chunks = (n_cols + 127) // 128 * 2  # 64-row units, padded to whole 128-column chunks
    Bp = np.zeros((n_chunks * CHUNK, 16), np.float32)
    Bp[:n_cols] = col_scale[:, None].astype(np.float32) * B.astype(np.float32)
    pieces = bf16_split(Bp)
    packed = np.zeros(n_chunks * CHUNK_BYTES // 2, np.uint16)
    j = np.arange(n_chunks * CHUNK)
    c, kl = j >> 6, j & 63
    base = c * CHUNK_BYTES + (kl >> 4) * KSTEP_BYTES + ((kl >> 3) & 1) * 768 + (kl & 7) * 2
    for p in range(3):
        for col in range(16):
            n = p * 16 + col
            off = base + (n >> 3) * 128 + (n & 7) * 16
            packed[off // 2] = pieces[p][:, col]
    return packed, Bp


def read_b_operand(packed, chunk, ks, chunk_cols=CHUNK):
    """what tcgen05.mma reads through bt_b_desc: element (n, k) of the 48 x 16 operand at
    start + (k / 8) * LBO + (n / 8) * SBO + (n % 8) * 16 + (k % 8) * 2 with LBO = 768, SBO = 128 (K-major, no swizzle);
    the image is chunk-size independent: k-step after k-step of 16 rows of B, a chunk is chunk_cols / 16 of them"""
    start = chunk * (chunk_cols // 16) * KSTEP_BYTES + ks * KSTEP_BYTES
    n = np.arange(48)[:, None]
    k = np.arange(16)[None, :]
    off = start + (k // 8) * 768 + (n // 8) * 128 + (n % 8) * 16 + (k % 8) * 2
    bits = packed[off // 2].astype(np.uint32) << np.uint32(16)
    return bits.view(np.float32)  # [48, 16]


def expand_words(words):
    """bt_expand_word on the 128 rows of a tile: 64-bit word -> 32 registers (two bf16 each) -> A[128, 64] of 0.0 / 1.0"""
    A = np.zeros((ROWS, CHUNK), np.float32)
    for half in range(2):
        w = ((words >> np.uint64(32 * half)) & np.uint64(0xFFFFFFFF)).astype(np.uint64)
        w8 = w >> np.uint64(8)
        for q in range(16):
            src = w if q < 8 else w8
            qq = q & 7
            reg = ((src & np.uint64(0x00010001 << qq)) * np.uint64(0x3F80 >> qq)) & np.uint64(0xFFFFFFFF)
            lo = ((reg & np.uint64(0xFFFF)) << np.uint64(16)).astype(np.uint32).view(np.float32)
            hi = (reg & np.uint64(0xFFFF0000)).astype(np.uint32).view(np.float32)
            col = half * 32 + 2 * q  # register q of the half = TMEM column (half*16 + q) = k elements 2q, 2q+1
            A[:, col] = lo
            A[:, col + 1] = hi
    assert np.all((A == 0) | (A == 1))
    return A


def emulate(plan, B):
    n_rows, CH, RB = plan["n_rows"], plan["chunk"], plan["rb"]
    BH = ROWS * RB
    NACC = 1 if RB == 2 else (4 if CH == 64 else 2)  # bt_mma_kernel / bt_mma_wide_kernel<1, 2> / <2, 1>
    packed, _ = pack_b(B, plan["col_scale"])
    P = np.zeros((plan["n_blk"] * BH, 16), np.float64)
    cells = []
    block_seen = np.zeros(plan["n_blk"], np.int64)
    tile_seen = np.zeros(plan["n_tiles"], np.int64)
    for q in range(plan["n_cta"]):
        t0, t1 = int(plan["cta_tile_ptr"][q]), int(plan["cta_tile_ptr"][q + 1])
        pos = 0
        for blk, end in plan["items"][plan["cta_item_ptr"][q]:plan["cta_item_ptr"][q + 1]]:
            blk, end = int(blk), int(end)
            assert end > pos, "an item owns at least one tile"
            block_seen[blk] += 1
            acc = np.zeros((RB, NACC, ROWS, 48), np.float64)
            chunks = plan["tile_chunk"][t0 + pos:t0 + end]
            assert np.all(np.diff(chunks.astype(np.int64)) > 0), "chunks of a block ascend"
            for idx in range(end - pos):
                t = t0 + pos + idx
                tile_seen[t] += 1
                for h in range(RB):  # the two 128-row halves of a 256-row item share the B' stage
                    A = np.concatenate([expand_words(plan["bits"][t][h][:, w]) for w in range(CH // 64)], 1)  # [128, CH]
                    for ks in range(CH // 16):
                        Bop = read_b_operand(packed, int(chunks[idx]), ks, CH)  # [48, 16]
                        acc[h, idx % NACC] += A[:, ks * 16:(ks + 1) * 16].astype(np.float64) @ Bop.T.astype(np.float64)
                    r, c = np.nonzero(A)
                    cells.append(np.stack([blk * BH + h * ROWS + r, int(chunks[idx]) * CH + c], 1))
            for h in range(RB):
                s = acc[h, 0].copy()
                for a in range(1, min(NACC, end - pos)):
                    s += acc[h, a]
                tot = (s[:, 32:48] + s[:, 16:32]) + s[:, 0:16]
                rows = np.arange(blk * BH + h * ROWS, blk * BH + (h + 1) * ROWS)
                sc = np.where(rows < n_rows, plan["row_scale"][np.minimum(rows, max(n_rows - 1, 0))], 0.0)
                P[rows] = sc[:, None] * tot
            pos = end
        assert t0 + pos == t1
    assert np.all(block_seen <= 1) and np.all(tile_seen == 1)
    rp = plan["r_indptr"].astype(np.int64)
    R = np.zeros((n_rows, 16), np.float64)
    rrows = np.repeat(np.arange(n_rows), np.diff(rp))
    np.add.at(R, rrows, plan["r_values"][:, None].astype(np.float64) * B[plan["r_indices"]])
    cells = np.concatenate(cells) if cells else np.zeros((0, 2), np.int64)
    cells = cells[cells[:, 0] < max(n_rows, 1)] if len(cells) else cells
    return P[:n_rows] + R, cells, rrows


def check_plan(gcnb, indptr, indices, values, B, explicit_scales=False, min_tile_nnz=0, n_cta=0, expect_tiles=True,
               chunk_cols=0, row_blocks=0):
    n = len(indptr) - 1
    deg = np.diff(indptr.astype(np.int64))
    rs = cs = None
    if explicit_scales:
        rs = cs = (1.0 / np.sqrt(deg.astype(np.float32))).astype(np.float32)
    plan = gcnb.bittile_host_build(indptr, indices, values, n, rs, cs, min_tile_nnz=min_tile_nnz, n_cta=n_cta,
                                   chunk_cols=chunk_cols, row_blocks=row_blocks)
    assert plan["chunk"] == (chunk_cols or 64) and plan["rb"] == (row_blocks or 1)
    assert plan["nnz"] == indices.size and plan["tile_nnz"] + plan["rem_nnz"] == indices.size
    out, cells, rrows = emulate(plan, B)
    assert len(cells) == plan["tile_nnz"]
    if expect_tiles:
        assert plan["n_tiles"] > 0 and plan["tile_nnz"] > 0
    # every CSR entry exactly once: multiset of (row, col) over bit maps + remainder == the CSR's
    rows = np.repeat(np.arange(n), deg)
    orig = np.sort(rows.astype(np.int64) * (1 << 32) + indices)
    got = np.sort(np.concatenate([cells[:, 0] * (1 << 32) + cells[:, 1],
                                  rrows.astype(np.int64) * (1 << 32) + plan["r_indices"]]))
    assert np.array_equal(orig, got)
    # remainder keeps the row's entry order and the original values
    ref = np.zeros((n, 16), np.float64)
    np.add.at(ref, rows, values[:, None].astype(np.float64) * B[indices])
    assert_close(out, ref, rtol=2e-6, atol=1e-7 * np.abs(ref).max(), what="bit-tile emulation vs CSR product")
    return plan


def test_bit_position_formula_is_a_permutation(gcnb):
    pos = [(c & 32) + ((c & 31) >> 1) + 16 * (c & 1) for c in range(64)]
    assert sorted(pos) == list(range(64))
    w = np.zeros(128, np.uint64)
    for c in range(64):
        w[c] |= np.uint64(1) << np.uint64(pos[c])  # row c has exactly column c set
    A = expand_words(w)
    assert np.array_equal(A[:64], np.eye(64, dtype=np.float32)) and not A[64:].any()


SHAPES = [(0, 0), (128, 1), (64, 2)]  # (columns per tile, 128-row blocks per item)


@pytest.mark.parametrize("chunk_cols,row_blocks", SHAPES)
def test_community_graph_matches_csr_product(gcnb, chunk_cols, row_blocks):
    rng = np.random.default_rng(7)
    indptr, indices, values = gcn_graph(rng, 1500, 5, 24, 3)
    B = rng.standard_normal((1500, 16)).astype(np.float32)
    thr = 96 * (2 if chunk_cols == 128 or row_blocks == 2 else 1)
    kw = dict(min_tile_nnz=thr, n_cta=5, chunk_cols=chunk_cols, row_blocks=row_blocks)
    plan = check_plan(gcnb, indptr, indices, values, B, **kw)
    assert plan["tile_nnz"] > 0.5 * indices.size
    # explicit 1/sqrt(deg) scales select the same entries
    plan2 = check_plan(gcnb, indptr, indices, values, B, explicit_scales=True, **kw)
    assert plan2["tile_nnz"] == plan["tile_nnz"]


@pytest.mark.parametrize("chunk_cols,row_blocks", SHAPES)
def test_duplicates_missing_diagonals_and_ragged_sizes(gcnb, chunk_cols, row_blocks):
    rng = np.random.default_rng(11)
    n = 777  # not a multiple of 128 or 64
    indptr, indices, values = gcn_graph(rng, n, 3, 30, 2, dup=40, drop_diag=(5, 300, 776))
    B = rng.standard_normal((n, 16)).astype(np.float32)
    plan = check_plan(gcnb, indptr, indices, values, B, min_tile_nnz=64, n_cta=148, chunk_cols=chunk_cols, row_blocks=row_blocks)
    # rows without a diagonal entry have no scale: none of their entries (nor entries pointing at them) is in a bit map
    for i in (5, 300, 776):
        assert plan["row_scale"][i] == 0 and plan["col_scale"][i] == 0
    # values that do not factor stay in the remainder with their original value
    values2 = values.copy()
    values2[::7] *= 1.5
    check_plan(gcnb, indptr, indices, values2, B, min_tile_nnz=64, n_cta=7, chunk_cols=chunk_cols, row_blocks=row_blocks)


def test_sparse_graph_has_no_tiles_and_empty_graph(gcnb):
    rng = np.random.default_rng(3)
    n = 4000
    deg = np.full(n, 3)
    indptr = np.zeros(n + 1, np.uint32)
    indptr[1:] = np.cumsum(deg)
    indices = rng.integers(0, n, int(indptr[-1])).astype(np.uint32)
    indices[indptr[:-1]] = np.arange(n)
    values = rng.random(indices.size).astype(np.float32)
    B = rng.standard_normal((n, 16)).astype(np.float32)
    plan = check_plan(gcnb, indptr, indices, values, B, expect_tiles=False)
    assert plan["n_tiles"] == 0 and plan["rem_nnz"] == indices.size
    empty = gcnb.bittile_host_build(np.zeros(1, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.float32), 0)
    assert empty["n_tiles"] == 0 and empty["n_rows"] == 0


def test_schedule_is_balanced_and_deterministic(gcnb):
    rng = np.random.default_rng(5)
    indptr, indices, values = gcn_graph(rng, 3000, 6, 20, 2)
    a = gcnb.bittile_host_build(indptr, indices, values, 3000, min_tile_nnz=64, n_cta=8, n_threads=1)
    b = gcnb.bittile_host_build(indptr, indices, values, 3000, min_tile_nnz=64, n_cta=8, n_threads=7)
    for k in ("tile_chunk", "bits", "cta_tile_ptr", "cta_item_ptr", "items", "r_indptr", "r_indices", "r_values"):
        assert np.array_equal(a[k], b[k]), k
    load = np.diff(a["cta_tile_ptr"].astype(np.int64))
    assert load.max() <= load.mean() * 1.5 + 8


# ---- mbarrier protocol of the MMA kernels, simulated -----------------------------------------------------------------
# A discrete-event model of one CTA of bt_mma_kernel / bt_mma_wide_kernel: the warp roles are state machines that follow
# the kernels' stage / use / parity arithmetic line by line, the barriers have the hardware's semantics (pending count,
# transaction bytes, phase parity -- a wait sees only the PARITY of the phase, so an overrun by two phases blocks for
# ever), tcgen05.commit arrivals and bulk-copy completions are delayed events that fire in issue order.  A random
# scheduler interleaves everything; the run must finish (no deadlock) and every MMA must find in its stage exactly the
# tile / chunk it is meant to consume.
class _Bar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.done = count, count, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.done += 1
            self.pending = self.count

    def arrive(self, tx=0):
        assert self.pending > 0, "more arrivals than the barrier expects in this phase"
        self.tx += tx
        self.pending -= 1
        self._check()

    def complete_tx(self, n):
        self.tx -= n
        self._check()

    def passed(self, parity):  # mbarrier.try_wait.parity
        return (self.done & 1) != parity


def _simulate_cta(items, unified, stages_a, stages_b, n_acc, rng):
    """items: list of tile counts per row block.  Returns the number of scheduler steps."""
    T = sum(items)
    ends = np.cumsum(items)
    if unified:
        full = [_Bar(5) for _ in range(stages_a)]
        free = [_Bar(1) for _ in range(stages_a)]
        a_full = b_full = full
        a_empty = b_empty = free
        stages_b = stages_a
    else:
        a_full, a_empty = [_Bar(4) for _ in range(stages_a)], [_Bar(1) for _ in range(stages_a)]
        b_full, b_empty = [_Bar(1) for _ in range(stages_b)], [_Bar(1) for _ in range(stages_b)]
    acc_full, acc_empty = [_Bar(1), _Bar(1)], [_Bar(4), _Bar(4)]
    A = [[None] * 4 for _ in range(stages_a)]   # tile held by (stage, lane quarter)
    B = [None] * stages_b
    acc_owner = [None, None]                    # item whose sums sit in the accumulator set
    events = []                                 # delayed arrivals, fired in order per queue
    copies = []

    def expander(g, quarter):
        for t in range(g, T, 2):
            s, use = t % stages_a, t // stages_a
            if use > 0:
                while not a_empty[s].passed((use - 1) & 1):
                    yield
            A[s][quarter] = t
            a_full[s].arrive()
            yield

    def producer():
        for t in range(T):
            s, use = t % stages_b, t // stages_b
            if use > 0:
                while not b_empty[s].passed((use - 1) & 1):
                    yield
            b_full[s].arrive(tx=6144)
            copies.append((s, t))
            yield

    def mma():
        t = 0
        for k, n_tiles in enumerate(items):
            st, use = k & 1, k >> 1
            if use > 0:
                while not acc_empty[st].passed((use - 1) & 1):
                    yield
            acc_owner[st] = k
            for idx in range(n_tiles):
                sa, ua, sb, ub = t % stages_a, t // stages_a, t % stages_b, t // stages_b
                while not a_full[sa].passed(ua & 1):
                    yield
                while not b_full[sb].passed(ub & 1):
                    yield
                assert A[sa] == [t] * 4, ("A stage holds", A[sa], "MMA wants", t)
                assert B[sb] == t, ("B stage holds", B[sb], "MMA wants", t)
                assert acc_owner[st] == k
                if unified:
                    events.append(free[sa])
                else:
                    events.append(a_empty[sa])
                    events.append(b_empty[sb])
                t += 1
                yield
            events.append(acc_full[st])
            yield

    def epilogue():
        for k in range(len(items)):
            st, use = k & 1, k >> 1
            while not acc_full[st].passed(use & 1):
                yield
            assert acc_owner[st] == k, "the accumulator set was overwritten before the epilogue read it"
            acc_empty[st].arrive()
            yield

    agents = [expander(g, q) for g in range(2) for q in range(4)] + [producer(), mma()] + [epilogue() for _ in range(4)]
    alive = list(range(len(agents)))
    steps = idle = 0
    while alive or events or copies:
        steps += 1
        assert steps < 400 * (T + len(items) + 10), "no progress: the protocol deadlocks"
        r = rng.random()
        if events and r < 0.15:
            events.pop(0).arrive()           # commits complete in issue order
            continue
        if copies and r < 0.3:
            s, t = copies.pop(int(rng.integers(0, len(copies))))  # bulk copies may land out of order
            B[s] = t
            b_full[s].complete_tx(6144)
            continue
        if not alive:
            continue
        i = alive[int(rng.integers(0, len(alive)))]
        try:
            next(agents[i])
        except StopIteration:
            alive.remove(i)
    return steps


@pytest.mark.parametrize("name,unified,stages_a,stages_b,n_acc", [
    ("bt_mma_kernel", False, 4, 8, 4), ("wide<1,1>", True, 8, 8, 2), ("wide<1,2>", True, 4, 4, 2), ("wide<2,1>", True, 5, 5, 1)])
def test_barrier_protocol_simulation(name, unified, stages_a, stages_b, n_acc):
    rng = np.random.default_rng(len(name))
    for items in ([1], [3], [4, 1, 1], [9, 2, 7, 1, 12], [37, 5, 1, 1, 2, 40], list(rng.integers(1, 30, 12))):
        for _ in range(3):
            _simulate_cta([int(x) for x in items], unified, stages_a, stages_b, n_acc, rng)


@pytest.mark.parametrize("chunk_cols,row_blocks", SHAPES)
def test_row_block_of_a_partitioned_graph(gcnb, chunk_cols, row_blocks):
    """a rank's row block of a row-partitioned GraphSum: rectangular (rows [r0, r1) x all columns), explicit global scales"""
    rng = np.random.default_rng(13)
    n, r0, r1 = 1500, 517, 1203
    indptr, indices, values = gcn_graph(rng, n, 5, 24, 3)
    s = (1.0 / np.sqrt(np.diff(indptr.astype(np.int64)).astype(np.float32))).astype(np.float32)
    ip = (indptr[r0:r1 + 1] - indptr[r0]).astype(np.uint32)
    ix, v = indices[indptr[r0]:indptr[r1]], values[indptr[r0]:indptr[r1]]
    B = rng.standard_normal((n, 16)).astype(np.float32)
    plan = gcnb.bittile_host_build(ip, ix, v, n, s[r0:r1], s, min_tile_nnz=48 * (2 if (chunk_cols == 128 or row_blocks == 2) else 1),
                                   n_cta=4, chunk_cols=chunk_cols, row_blocks=row_blocks)
    assert plan["n_rows"] == r1 - r0 and plan["n_cols"] == n and plan["n_tiles"] > 0
    out, cells, rrows = emulate(plan, B)
    rows = np.repeat(np.arange(r1 - r0), np.diff(ip.astype(np.int64)))
    ref = np.zeros((r1 - r0, 16), np.float64)
    np.add.at(ref, rows, v[:, None].astype(np.float64) * B[ix])
    assert_close(out, ref, rtol=2e-6, atol=1e-7 * np.abs(ref).max(), what="row block")
    got = np.sort(np.concatenate([cells[:, 0] * (1 << 32) + cells[:, 1], rrows.astype(np.int64) * (1 << 32) + plan["r_indices"]]))
    assert np.array_equal(got, np.sort(rows.astype(np.int64) * (1 << 32) + ix))
    # without explicit scales a rectangular block has no diagonal to derive them from: everything stays in the remainder
    none = gcnb.bittile_host_build(ip, ix, v, n, chunk_cols=chunk_cols, row_blocks=row_blocks)
    assert none["tile_nnz"] == 0 and none["rem_nnz"] == len(ix)
