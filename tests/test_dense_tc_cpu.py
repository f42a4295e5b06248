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
    assert gcnb.lib.gcnb_dense_tc_w_bytes(602, 600) == 4 * ks * 3 * 160 * 32  # 600 columns -> 4 parts of 160 (3 accumulator classes)
    assert gcnb.lib.gcnb_dense_tc_w_bytes(50, 16) == 1 * 4 * 3 * 16 * 32


# ---- operand images: the pack kernels' write offsets against the descriptors' read addressing (dense_tc.cu restated) ------
def _bf16_bits(x):
    return (x.astype(np.float32).view(np.uint32) >> np.uint32(16)).astype(np.uint16)


def _from_bits(b):
    return (b.astype(np.uint32) << np.uint32(16)).view(np.float32).astype(np.float64)


def _shape(f, p):
    KS = (f + 15) // 16
    p_pad = (p + 15) // 16 * 16
    n_parts = (p_pad + 159) // 160
    pcols = ((p_pad + n_parts - 1) // n_parts + 15) // 16 * 16
    return KS, n_parts, pcols


def _pack_rows(M, n_rows_pad, KS):
    """tc_pack_x_kernel / tc_pack_xt_kernel: M[row][k] (row = operand row) -> tiles ((b * KS + ks) * 3 + piece) of 4096 bytes,
    element at half * 2048 + (r / 8) * 128 + (r % 8) * 16 + (k % 8) * 2"""
    rows, K = M.shape
    n_blk = n_rows_pad // 128
    img = np.zeros(n_blk * KS * 3 * 2048, np.uint16)
    Mp = np.zeros((n_rows_pad, KS * 16), np.float32)
    Mp[:rows, :K] = M
    pieces = [_bf16_bits(p) for p in split3(Mp)]
    r = np.arange(n_rows_pad)[:, None]
    k = np.arange(KS * 16)[None, :]
    b, rr, ks, kk = r // 128, r % 128, k // 16, k % 16
    for pc in range(3):
        off = ((b * KS + ks) * 3 + pc) * 4096 + (kk // 8) * 2048 + (rr // 8) * 128 + (rr % 8) * 16 + (kk % 8) * 2
        img[off // 2] = pieces[pc]
    return img


def _pack_w(W, f, p):
    """tc_pack_w_kernel: tile ((q * KS + ks) * 3 + piece) of pcols * 32 bytes, element (column c, k) at
    (k / 8) * pcols * 16 + (c / 8) * 128 + (c % 8) * 16 + (k % 8) * 2"""
    KS, n_parts, pcols = _shape(f, p)
    Wp = np.zeros((KS * 16, n_parts * pcols), np.float32)
    Wp[:f, :p] = W
    pieces = [_bf16_bits(x) for x in split3(Wp)]
    img = np.zeros(n_parts * KS * 3 * pcols * 16, np.uint16)
    k = np.arange(KS * 16)[:, None]
    col = np.arange(n_parts * pcols)[None, :]
    q, c, ks, kk = col // pcols, col % pcols, k // 16, k % 16
    for pc in range(3):
        off = ((q * KS + ks) * 3 + pc) * (pcols * 32) + (kk // 8) * (pcols * 16) + (c // 8) * 128 + (c % 8) * 16 + (kk % 8) * 2
        img[off // 2] = pieces[pc]
    return img


def _read_operand(img, base_bytes, rows, lbo):
    """what tcgen05.mma reads through tcg_desc(base, lbo): element (row, k) at base + (k / 8) * lbo + (row / 8) * 128 +
    (row % 8) * 16 + (k % 8) * 2"""
    r = np.arange(rows)[:, None]
    k = np.arange(16)[None, :]
    off = base_bytes + (k // 8) * lbo + (r // 8) * 128 + (r % 8) * 16 + (k % 8) * 2
    return _from_bits(img[off // 2])


def _gemm_emulated(a_img, b_img, n_rows_pad, KS, n_parts, pcols, k_slices=1):
    """tc_gemm_kernel: per (row block, part, k slice) the six piece products per k-step, slices added in ascending order"""
    n_blk = n_rows_pad // 128
    out = np.zeros((n_rows_pad, n_parts * pcols))
    for blk in range(n_blk):
        for q in range(n_parts):
            for sl in range(k_slices):
                acc = np.zeros((128, pcols))
                for ks in range(KS * sl // k_slices, KS * (sl + 1) // k_slices):
                    A = [_read_operand(a_img, ((blk * KS + ks) * 3 + pc) * 4096, 128, 2048) for pc in range(3)]
                    B = [_read_operand(b_img, ((q * KS + ks) * 3 + pc) * pcols * 32, pcols, pcols * 16) for pc in range(3)]
                    for i, j in ((2, 0), (0, 2), (1, 1), (1, 0), (0, 1), (0, 0)):
                        acc += A[i] @ B[j].T
                out[blk * 128:(blk + 1) * 128, q * pcols:(q + 1) * pcols] += acc
    return out


def test_forward_operand_images_and_descriptor_addressing():
    rng = np.random.default_rng(1)
    for n, f, p in ((200, 50, 41), (130, 602, 600)):
        X = rng.standard_normal((n, f)).astype(np.float32)
        W = rng.standard_normal((f, p)).astype(np.float32)
        KS, n_parts, pcols = _shape(f, p)
        n_pad = (n + 127) // 128 * 128
        got = _gemm_emulated(_pack_rows(X, n_pad, KS), _pack_w(W, f, p), n_pad, KS, n_parts, pcols)[:n, :p]
        assert_close(got, X.astype(np.float64) @ W.astype(np.float64), what="forward image %dx%dx%d" % (n, f, p))


def test_weight_gradient_images_and_split_k():
    rng = np.random.default_rng(2)
    n, f, p = 1000, 150, 41  # operand rows = features (2 blocks of 128), K = nodes (63 k-steps), 5 k slices
    X = rng.standard_normal((n, f)).astype(np.float32)
    dH = rng.standard_normal((n, p)).astype(np.float32)
    KS, n_parts, pcols = _shape(n, p)
    f_pad = (f + 127) // 128 * 128
    a_img = _pack_rows(np.ascontiguousarray(X.T), f_pad, KS)  # tc_pack_xt_kernel: operand row = feature, k = node
    b_img = _pack_w(dH, n, p)                                 # tc_pack_w_kernel with K = n
    want = X.astype(np.float64).T @ dH.astype(np.float64)
    for k_slices in (1, 5, 63):
        got = _gemm_emulated(a_img, b_img, f_pad, KS, n_parts, pcols, k_slices)[:f, :p]
        assert_close(got, want, what="X^T dH, %d k slices" % k_slices)


# ---- mbarrier protocol of tc_gemm_kernel, simulated (same model as tests/test_bittile_cpu.py: parity-only waits, delayed
# ---- commits in issue order, bulk copies landing out of order, random interleaving of the roles)
def test_gemm_barrier_protocol_simulation():
    from tests.test_bittile_cpu import _Bar
    rng = np.random.default_rng(4)
    stages = 4
    for items in ([1], [3, 1], [38, 38, 38], [5, 9, 2, 7, 1, 1, 12], list(rng.integers(1, 20, 9))):
        for _ in range(3):
            full, free = [_Bar(1) for _ in range(stages)], [_Bar(1) for _ in range(stages)]
            acc_full, acc_empty = [_Bar(1), _Bar(1)], [_Bar(4), _Bar(4)]
            stage, acc_owner, events, copies = [None] * stages, [None, None], [], []
            T = int(sum(items))

            def producer():
                for t in range(T):
                    s, use = t % stages, t // stages
                    if use > 0:
                        while not free[s].passed((use - 1) & 1):
                            yield
                    full[s].arrive(tx=2)          # arrive.expect_tx, then two bulk copies (A tiles, B tiles)
                    copies.append((s, t, "a"))
                    copies.append((s, t, "b"))
                    yield

            def mma():
                t = 0
                for k, n in enumerate(items):
                    st, use = k & 1, k >> 1
                    if use > 0:
                        while not acc_empty[st].passed((use - 1) & 1):
                            yield
                    acc_owner[st] = k
                    for _ in range(int(n)):
                        s, u = t % stages, t // stages
                        while not full[s].passed(u & 1):
                            yield
                        assert stage[s] == {"a": t, "b": t}, (stage[s], t)
                        events.append(free[s])
                        t += 1
                        yield
                    events.append(acc_full[st])
                    yield

            def epilogue():
                for k in range(len(items)):
                    st, use = k & 1, k >> 1
                    while not acc_full[st].passed(use & 1):
                        yield
                    assert acc_owner[st] == k
                    acc_empty[st].arrive()
                    yield

            agents = [producer(), mma()] + [epilogue() for _ in range(4)]
            alive = list(range(len(agents)))
            steps = 0
            while alive or events or copies:
                steps += 1
                assert steps < 2000 * (T + 10), "deadlock"
                r = rng.random()
                if events and r < 0.15:
                    events.pop(0).arrive()
                    continue
                if copies and r < 0.35:
                    s, t, which = copies.pop(int(rng.integers(0, len(copies))))
                    if stage[s] is None or stage[s].get(which) is not None and stage[s].get("a") == stage[s].get("b") and stage[s]["a"] != t:
                        stage[s] = {}
                    stage[s] = dict(stage[s] or {}, **{which: t})
                    full[s].complete_tx(1)
                    continue
                if not alive:
                    continue
                i = alive[int(rng.integers(0, len(alive)))]
                try:
                    next(agents[i])
                except StopIteration:
                    alive.remove(i)
