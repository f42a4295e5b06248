"""CPU-only tests (-m "not gpu"): the oracle against the reference's own CPU code and the golden fixtures, the host
logic (parser, synthetic generator, RNG bookkeeping) and the C ABI surface (every symbol of include/*.h is exported;
no compute call is made without a GPU)."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32, u32, i32, u8 = np.float32, np.uint32, np.int32, np.uint8


@pytest.fixture(scope="module")
def eng():
    import __graft_entry__ as ge
    ge.load_package()
    import importlib
    return importlib.import_module("parallel_gcn_b200.engine")


def test_abi_exports_every_declared_symbol():
    declared = set()
    for h in ("gcnb.h", "gcnb_engine.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        declared |= set(re.findall(r"GCNB_API\s+[\w\s\*]+?\b(gcnb_\w+)\s*\(", src))
    assert len(declared) > 40
    lib = C.CDLL(os.path.join(ROOT, "parallel-gcn_b200", "libgcn_b200.so"))
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    lib.gcnb_error_string.restype = C.c_char_p
    assert lib.gcnb_version() >= 100 and b"bad argument" in lib.gcnb_error_string(10001)


def test_no_gpu_means_error_not_fallback(eng):
    """the product path must fail loudly without a device (never route through a CPU implementation)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ds = eng.parse_dataset(ROOT, "cora")
    with pytest.raises(eng.GcnbError):
        eng.GCN(ds)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "parallel-gcn_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("oracle/", "oracle/") or "import oracle" not in txt, fn
                assert "liboracle" not in txt and "libref_cpu" not in txt and "from oracle" not in txt, fn


@pytest.mark.parametrize("name", ["cora", "citeseer"])
def test_parser_bit_exact(O, eng, name):
    """product Parser == oracle parser == the reference's own Parser (when oracle/_ref is built): CSR arrays, labels,
    split, dims, split counts and graph_value bit for bit."""
    mine = eng.parse_dataset(ROOT, name)
    orc = O.parse_dataset(os.path.join(ROOT, "data", name))
    for k in ("g_indptr", "g_indices", "f_indptr", "f_indices", "label", "split"):
        assert (getattr(mine, k) == getattr(orc, k)).all(), k
    assert (mine.f_value.view(u32) == orc.f_value.view(u32)).all()
    assert (mine.graph_value.view(u32) == orc.graph_values().view(u32)).all()
    assert (mine.input_dim, mine.output_dim) == (orc.input_dim, orc.output_dim)
    assert mine.split_counts == orc.split_counts()
    if O.ref is not None:
        _, rds = O.ref_parse_dataset(ROOT, name)
        for k in ("g_indptr", "g_indices", "f_indptr", "f_indices", "label", "split"):
            assert (getattr(mine, k).view(i32) == getattr(rds, k)).all(), k
        assert (mine.f_value.view(u32) == rds.f_value.view(u32)).all()
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "datasets.json")))[name]
    assert gold["num_nodes"] == mine.num_nodes and gold["graph_nnz"] == len(mine.g_indices)
    assert gold["feat_nnz"] == len(mine.f_indices) and gold["split_counts"] == list(mine.split_counts)
    for k in ("g_indptr", "g_indices", "f_indptr", "f_indices", "f_value", "label", "split", "graph_value"):
        assert gold["fnv"][k] == O.fnv(getattr(mine, k)), k


def test_parser_pubmed_graph_only(O, eng):
    """pubmed ships without .svmlight: the parser must report failure exactly like the reference (parse() == false)."""
    assert eng.parse_dataset(ROOT, "pubmed") is None
    assert O.parse_dataset(os.path.join(ROOT, "data", "pubmed")) is None


def test_parser_edge_cases(O, eng, tmp_path):
    """ragged input: empty neighbour lines, blank svmlight lines (label -1), duplicate neighbours, explicit self entries,
    CRLF, unterminated last line (dropped by the reference's getline/eof loop), NO_FEATURE mode."""
    d = tmp_path / "data"
    d.mkdir()
    (d / "t.graph").write_text("1 2 2\n\n0 0\r\n3\n2")
    (d / "t.svmlight").write_text("1 0:0.5 3:1e-3\n\n0 2:+2.5 2:7\n2 5:-1\n1 1:1")
    (d / "t.split").write_text("1\n2\n3\n1\n1")
    mine = eng.parse_dataset(tmp_path, "t")
    orc = O.parse_dataset(str(d / "t"))
    assert mine.num_nodes == 4 and list(mine.g_indptr) == [0, 4, 5, 8, 10]
    assert list(mine.g_indices) == [0, 1, 2, 2, 1, 2, 0, 0, 3, 3]
    assert list(mine.label) == [1, -1, 0, 2] and list(mine.f_indptr) == [0, 2, 2, 4, 5]
    assert mine.input_dim == 6 and mine.output_dim == 3 and mine.split_counts == (2, 1, 1)
    for k in ("g_indptr", "g_indices", "f_indptr", "f_indices", "label", "split"):
        assert (getattr(mine, k) == getattr(orc, k)).all(), k
    assert (mine.f_value.view(u32) == orc.f_value.view(u32)).all()
    if O.ref is not None:
        _, rds = O.ref_parse_dataset(str(tmp_path), "t")
        for k in ("g_indptr", "g_indices", "f_indptr", "f_indices", "label", "split"):
            assert (getattr(mine, k).view(i32) == getattr(rds, k)).all(), k
        assert (mine.f_value.view(u32) == rds.f_value.view(u32)).all()
    nf = eng.parse_dataset(tmp_path, "t", no_feature=True)
    assert (nf.f_value == 1.0).all()


# ---- randomised parser parity: product parser == oracle parser == the reference's OWN parser on ragged files ----
def fmt_float(rng, x):
    k = rng.integers(0, 7)
    if k == 0: return "%g" % x
    if k == 1: return "%.6f" % x
    if k == 2: return "%e" % x
    if k == 3: return ("+%g" % abs(x))
    if k == 4: return "%d" % int(round(x))
    if k == 5: return ("%.3f" % x).lstrip("0") if 0 < x < 1 else "%g" % x   # ".5"
    return "%d." % int(round(x))                                               # "3."

def make(rng, d, name, n):
    sep = lambda: " " * int(rng.integers(1, 3))
    eol = lambda: "\r\n" if rng.random() < 0.1 else "\n"
    g, s, v = [], [], []
    for i in range(n):
        deg = int(rng.integers(0, 6))
        nb = [str(int(rng.integers(0, max(1, n - 1)))) for _ in range(deg)]  # valid even if the last line is dropped
        if rng.random() < 0.2 and nb: nb.append(nb[0])          # duplicate neighbour
        if rng.random() < 0.2 and i < n - 1: nb.append(str(i))  # explicit self entry
        g.append(sep().join(nb) + (" " if rng.random() < 0.2 else "") + eol())
        if rng.random() < 0.1:
            v.append(eol())                                     # blank svmlight line: label -1
        else:
            k = int(rng.integers(0, 5))
            idx = sorted(set(int(x) for x in rng.integers(0, 12, k)))
            feats = ["%d:%s" % (j, fmt_float(rng, float(rng.normal()) * 10 ** int(rng.integers(-3, 3)))) for j in idx]
            v.append(sep().join([str(int(rng.integers(0, 4)))] + feats) + eol())
        s.append(str(int(rng.integers(1, 4))) + eol())
    if rng.random() < 0.3:                                      # unterminated last line
        g[-1] = g[-1].rstrip("\r\n"); v[-1] = v[-1].rstrip("\r\n"); s[-1] = s[-1].rstrip("\r\n")
    open(os.path.join(d, name + ".graph"), "w", newline="").write("".join(g))
    open(os.path.join(d, name + ".svmlight"), "w", newline="").write("".join(v))
    open(os.path.join(d, name + ".split"), "w", newline="").write("".join(s))



def test_parser_matches_reference_on_random_ragged_files(O, eng, tmp_path):
    """60 seeded random datasets with the quirks real files have: empty neighbour lists, duplicate neighbours, explicit self
    entries, blank svmlight lines (label -1), CRLF line ends, trailing blanks, unterminated last lines, numbers written as
    1e-3 / +2.5 / .5 / 3. / integers.  Every array must equal the reference parser's bit for bit (and the oracle's)."""
    for seed in range(60):
        rng = np.random.default_rng(seed)
        root = tmp_path / ("s%d" % seed)
        d = root / "data"
        d.mkdir(parents=True)
        make(rng, str(d), "t", int(rng.integers(1, 12)))
        mine = eng.parse_dataset(str(root), "t")
        orc = O.parse_dataset(str(d / "t"))
        assert (mine is None) == (orc is None), seed
        if mine is None:
            continue
        for k in ("g_indptr", "g_indices", "f_indptr", "f_indices", "label", "split"):
            assert (getattr(mine, k) == getattr(orc, k)).all(), (seed, k)
        assert (mine.f_value.view(u32) == orc.f_value.view(u32)).all(), seed
        assert (mine.input_dim, mine.output_dim) == (orc.input_dim, orc.output_dim), seed
        if O.ref is not None:
            _, rds = O.ref_parse_dataset(str(root), "t")
            assert rds is not None, seed
            for k in ("g_indptr", "g_indices", "f_indptr", "f_indices", "label", "split"):
                a, b = getattr(mine, k), getattr(rds, k)
                assert a.shape == b.shape and (a.view(i32) == b).all(), (seed, k)
            assert (mine.f_value.view(u32) == rds.f_value.view(u32)).all(), seed


def test_oracle_pinned_to_reference_cpu(O, datasets):
    """the C restatement reproduces the reference's own CPU implementation bit for bit: 3 epochs of losses, accuracies
    and both weight matrices (needs oracle/_ref, i.e. /root/reference at build time)."""
    if O.ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    for name in ("cora", "citeseer"):
        ds = datasets[name]
        h, _ = O.ref_parse_dataset(ROOT, name)
        O.ref.ref_srand(1)
        g = O.ref.ref_gcn_create(h, 16, 0.5, 0.01, 5e-4, 100, 0)
        O.lib.orc_libc_srand(1)
        og = O.OracleGCN(ds, flavour="ref_cpu")
        out = np.zeros(2, f32)
        for _ in range(3):
            O.ref.ref_gcn_train_epoch(g, O._p(out)); rt = tuple(out)
            O.ref.ref_gcn_eval(g, 2, O._p(out)); rv = tuple(out)
            ot, ov = og.train_epoch(), og.eval(2)
            assert (f32(ot[0]), f32(ot[1]), f32(ov[0]), f32(ov[1])) == (rt[0], rt[1], rv[0], rv[1])
        for idx, l in ((2, 0), (5, 1)):
            w = np.empty(og.W[l].size, f32)
            O.ref.ref_gcn_variable_get(g, idx, 0, O._p(w))
            assert (w.view(u32) == og.W[l].view(u32)).all()
        O.ref.ref_gcn_free(g)
        O.ref.ref_dataset_free(h)


def test_oracle_pinned_to_reference_cpu_on_random_ragged_datasets(O, tmp_path):
    """beyond cora / citeseer: 30 seeded random datasets (directed graphs with duplicate and self entries, isolated nodes,
    unlabelled nodes, rows without features, feature magnitudes over five decades) trained for 3 epochs by the reference's
    own CPU code and by the restatement -- every loss, accuracy and weight bit must agree."""
    if O.ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    trained = 0
    for seed in range(30):
        rng = np.random.default_rng(1000 + seed)
        root = tmp_path / ("r%d" % seed)
        d = root / "data"
        d.mkdir(parents=True)
        make(rng, str(d), "t", int(rng.integers(6, 40)))
        ds = O.parse_dataset(str(d / "t"))
        res = O.ref_parse_dataset(str(root), "t")
        if ds is None or res is None or res[1] is None:
            continue
        h = res[0]
        if not all(((ds.split == k) & (ds.label >= 0)).sum() > 0 for k in (1, 2)):
            O.ref.ref_dataset_free(h)
            continue
        trained += 1
        O.ref.ref_srand(1)
        g = O.ref.ref_gcn_create(h, 16, 0.5, 0.01, 5e-4, 100, 0)
        O.lib.orc_libc_srand(1)
        og = O.OracleGCN(ds, flavour="ref_cpu")
        out = np.zeros(2, f32)
        for ep in range(3):
            O.ref.ref_gcn_train_epoch(g, O._p(out)); rt = tuple(out)
            O.ref.ref_gcn_eval(g, 2, O._p(out)); rv = tuple(out)
            ot, ov = og.train_epoch(), og.eval(2)
            a, b = np.array([ot[0], ot[1], ov[0], ov[1]], f32), np.array([rt[0], rt[1], rv[0], rv[1]], f32)
            assert (a.view(u32) == b.view(u32)).all(), (seed, ep, a, b)
        for idx, l in ((2, 0), (5, 1)):
            w = np.empty(og.W[l].size, f32)
            O.ref.ref_gcn_variable_get(g, idx, 0, O._p(w))
            assert (w.view(u32) == og.W[l].view(u32)).all(), (seed, l)
        O.ref.ref_gcn_free(g)
        O.ref.ref_dataset_free(h)
    assert trained >= 20


def test_oracle_against_golden_training_curves(O, datasets):
    """committed fixtures generated from the reference's CPU code (tests/golden/make_golden.py)."""
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_cpu_training.json")))
    for name in ("cora", "citeseer"):
        O.lib.orc_libc_srand(1)
        og = O.OracleGCN(datasets[name], flavour="ref_cpu")
        for ep, want in enumerate(gold[name]["epochs"]):
            t, v = og.train_epoch(), og.eval(2)
            got = [t[0], t[1], v[0], v[1]]
            assert [float(f32(x)) for x in got] == want, (name, ep)
        assert O.fnv(og.W[0]) == gold[name]["w0_fnv"] and O.fnv(og.W[1]) == gold[name]["w1_fnv"]


def test_oracle_modules_against_reference(O):
    """module-level pin on random ragged inputs (empty rows, duplicates) against the reference's Module classes."""
    if O.ref is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(11)
    n, dim = 300, 9
    deg = rng.integers(0, 12, n); deg[0] = 0
    indptr = np.zeros(n + 1, i32); indptr[1:] = np.cumsum(deg)
    # the reference GraphSum needs every row non-empty (coef = 1/sqrt(deg*deg)): add self loops like the parser
    indices = np.concatenate([np.r_[i, rng.integers(0, n, d)] for i, d in enumerate(deg)]).astype(i32)
    indptr = np.zeros(n + 1, i32); indptr[1:] = np.cumsum(deg + 1)
    x = rng.standard_normal((n, dim)).astype(f32)
    og = rng.standard_normal((n, dim)).astype(f32)
    r_out, r_ig = np.empty_like(x), np.empty_like(x)
    O.ref.ref_graphsum(n, dim, O._p(indptr), O._p(indices), O._p(x), O._p(r_out), O._p(og), O._p(r_ig))
    o_out, o_ig = np.empty_like(x), np.empty_like(x)
    ip, ix = indptr.astype(u32), indices.astype(u32)
    O.lib.orc_graphsum(n, dim, O._p(ip), O._p(ix), None, O._p(x), O._p(o_out))
    O.lib.orc_graphsum(n, dim, O._p(ip), O._p(ix), None, O._p(og), O._p(o_ig))
    assert (r_out.view(u32) == o_out.view(u32)).all() and (r_ig.view(u32) == o_ig.view(u32)).all()
    # SparseMatmul / Matmul
    m, nn, p = 200, 37, 5
    fdeg = rng.integers(0, 6, m)
    fip = np.zeros(m + 1, i32); fip[1:] = np.cumsum(fdeg)
    fix = rng.integers(0, nn, fip[-1]).astype(i32)
    fv = rng.standard_normal(fip[-1]).astype(f32)
    b = rng.standard_normal((nn, p)).astype(f32)
    cg = rng.standard_normal((m, p)).astype(f32)
    rc, rbg = np.empty((m, p), f32), np.empty((nn, p), f32)
    O.ref.ref_sparse_matmul(m, nn, p, O._p(fip), O._p(fix), O._p(fv), O._p(b), O._p(rc), O._p(cg), O._p(rbg))
    oc, obg = np.empty((m, p), f32), np.empty((nn, p), f32)
    O.lib.orc_spmm(m, p, O._p(fip.astype(u32)), O._p(fix.astype(u32)), O._p(fv), O._p(b), O._p(oc))
    O.lib.orc_spmm_bwd(m, nn, p, O._p(fip.astype(u32)), O._p(fix.astype(u32)), O._p(fv), O._p(cg), O._p(obg))
    assert (rc.view(u32) == oc.view(u32)).all() and (rbg.view(u32) == obg.view(u32)).all()
    a = rng.standard_normal((m, nn)).astype(f32)
    rc, rag, rbg = np.empty((m, p), f32), np.empty((m, nn), f32), np.empty((nn, p), f32)
    O.ref.ref_matmul(m, nn, p, O._p(a), O._p(b), O._p(rc), O._p(cg), O._p(rag), O._p(rbg))
    oc, oag, obg = np.empty((m, p), f32), np.empty((m, nn), f32), np.empty((nn, p), f32)
    O.lib.orc_matmul(m, nn, p, O._p(a), O._p(b), O._p(oc))
    O.lib.orc_matmul_bwd(m, nn, p, O._p(a), O._p(b), O._p(cg), O._p(oag), O._p(obg))
    assert (rc.view(u32) == oc.view(u32)).all() and (rag.view(u32) == oag.view(u32)).all() and (rbg.view(u32) == obg.view(u32)).all()
    # CrossEntropy (ref-CPU flavour), ReLU, Adam, Glorot/Dropout streams
    C_ = 7
    lg = (rng.standard_normal((m, C_)) * 2).astype(f32)
    truth = rng.integers(-1, C_, m).astype(i32)
    rl, rgr = lg.copy(), np.empty((m, C_), f32)
    rloss = O.ref.ref_cross_entropy(m, C_, O._p(rl), O._p(truth), O._p(rgr), 1)
    ol, ogr = lg.copy(), np.empty((m, C_), f32)
    oloss = O.lib.orc_cross_entropy(m, C_, O._p(ol), O._p(truth), O._p(ogr), 0, 1, None)
    assert f32(rloss) == f32(oloss) and (rl.view(u32) == ol.view(u32)).all() and (rgr.view(u32) == ogr.view(u32)).all()
    O.ref.ref_rand_state_set(12345, 67890); O.lib.orc_xorshift_set(12345, 67890)
    rw, ow = np.empty(500, f32), np.empty(500, f32)
    O.ref.ref_glorot(500, 100, 5, O._p(rw)); O.lib.orc_glorot_xorshift(500, 100, 5, O._p(ow))
    assert (rw.view(u32) == ow.view(u32)).all()
    xr = np.ones(1000, f32); gr = np.empty(1000, f32)
    O.ref.ref_dropout(1000, 0.3, O._p(xr), O._p(gr))
    mk = np.empty(1000, u8); O.lib.orc_dropout_mask_xorshift(1000, 0.3, O._p(mk))
    xo = np.ones(1000, f32); O.lib.orc_dropout_apply(1000, O._p(xo), O._p(mk), O.lib.orc_dropout_scale(0.3, 0))
    assert (xr.view(u32) == xo.view(u32)).all() and ((gr != 0) == (mk != 0)).all()
    w0 = rng.standard_normal(300).astype(f32); grads = (rng.standard_normal((4, 300)) * 0.1).astype(f32)
    rw = w0.copy(); O.ref.ref_adam(300, 4, O._p(rw), O._p(grads), 1, 0.01, 5e-4)
    ow, mm, vv = w0.copy(), np.zeros(300, f32), np.zeros(300, f32)
    for s in range(4):
        ss = O.lib.orc_adam_step_size(0.01, 0.9, 0.999, s + 1)
        O.lib.orc_adam_step(300, O._p(ow), O._p(grads[s]), O._p(mm), O._p(vv), 1, 5e-4, 0.9, 0.999, 1e-8, ss)
    assert (rw.view(u32) == ow.view(u32)).all()


def test_philox_matches_published_vectors_and_curand_layout(O):
    """Philox4x32-10 known-answer tests (Random123 kat_vectors) + the cuRAND uniform mapping recorded in SURVEY 5.9."""
    def ph(ctr, key):
        out = np.zeros(4, u32)
        O.lib.orc_philox4x32_10(O._p(np.array(ctr, u32)), O._p(np.array(key, u32)), O._p(out))
        return [int(x) for x in out]
    assert ph([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert ph([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert ph([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = np.zeros(4, f32)
    O.lib.orc_curand_uniform4(19990304, 5, 0, 1, O._p(u))
    assert abs(float(u[0]) - 0.0917690247) < 1e-9
    O.lib.orc_curand_uniform4(19990304, 5, 1, 1, O._p(u))
    assert abs(float(u[3]) - 0.936051846) < 1e-9
    # the real cuRAND headers executed on the host (tests/native/curand_ref.cu -DCURAND_REF_HOST): every state/draw
    from tests.native import build as nb
    _, host = nb.build()
    lib = C.CDLL(host)
    n_states, n_draws, seed = 300, 4, 311288059
    ref = np.zeros((n_draws, n_states, 4), f32)
    assert lib.curand_ref_host(C.c_uint(seed), n_states, n_draws, ref.ctypes.data_as(C.c_void_p)) == 0
    for t in range(n_draws):
        for i in range(0, n_states, 7):
            O.lib.orc_curand_uniform4(seed, i, t, 0, O._p(u))
            assert (u.view(u32) == ref[t, i].view(u32)).all(), (t, i)


def test_synth_graph_properties(eng):
    """generator: exact edge count, symmetric, simple, self entry first, sorted rows, deterministic, community locality."""
    n, m = 5000, 60000
    ip, ix = eng.synth_graph(n, m, n_blocks=10, intra=0.8, sigma=1.0, max_deg=400, seed=7)
    ip2, ix2 = eng.synth_graph(n, m, n_blocks=10, intra=0.8, sigma=1.0, max_deg=400, seed=7)
    assert (ip == ip2).all() and (ix == ix2).all()
    assert len(ix) == 2 * m + n and ip[-1] == len(ix)
    rows = np.repeat(np.arange(n), np.diff(ip.astype(np.int64)))
    assert (ix[ip[:-1]] == np.arange(n)).all()           # implicit self index first (parser convention)
    off = np.ones(len(ix), bool); off[ip[:-1]] = False
    r, c = rows[off], ix[off].astype(np.int64)
    assert (r != c).all()
    key = r * n + c
    assert len(np.unique(key)) == len(key)                 # simple graph
    assert np.array_equal(np.sort(key), np.sort(c * n + r))  # symmetric
    assert (np.diff(key) > 0).all()                        # rows sorted ascending
    bs = (n + 9) // 10
    assert 0.7 < ((r // bs) == (c // bs)).mean() < 0.9
    ds = eng.synth_dataset(300, 2000, 12, 5, n_blocks=3, seed=3)
    assert ds.f_value.shape == (3600,) and abs(float(ds.f_value.mean())) < 0.1 and sum(ds.split_counts) == 300
    assert (ds.f_indices.reshape(300, 12) == np.arange(12)).all() and set(np.unique(ds.label)) <= set(range(5))


def test_reference_main_compiles_against_product_headers(tmp_path):
    """drop-in at source level: the reference's own src/main.cpp and its three GPU drivers (test/performance_gpu.cpp,
    tuning_cuda.cpp, tuning_accuracy.cpp, with the flags of the reference Makefile's targets) compile UNCHANGED against
    parallel-gcn_b200/host/include and link against libgcn_b200.so (needs /root/reference; compile+link only)."""
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("reference not present on this box")
    inc = os.path.join(ROOT, "parallel-gcn_b200", "host", "include")
    os.symlink(inc, tmp_path / "include")
    for sub, fn, flags in (("src", "main.cpp", []), ("src", "main.cpp", ["-DPART2", "-DNO_FEATURE"]),
                           ("test", "performance_gpu.cpp", ["-DNO_OUTPUT", "-DTUNE_CUDA", "-DPERFORMANCE"]),
                           ("test", "tuning_cuda.cpp", ["-DDEBUG_CUDA", "-DNO_OUTPUT", "-DTUNE_CUDA"]),
                           ("test", "tuning_accuracy.cpp", ["-DDEBUG_CUDA", "-DNO_OUTPUT", "-DTUNE_ACCURACY", "-DPART2"]),
                           ("test", "tuning_accuracy.cpp", ["-DDEBUG_CUDA", "-DNO_OUTPUT", "-DTUNE_ACCURACY", "-DNO_FEATURE", "-DPART2"])):
        (tmp_path / sub).mkdir(exist_ok=True)
        dst = tmp_path / sub / fn
        if not dst.exists():
            os.symlink(os.path.join(ref, sub, fn), dst)
        exe = tmp_path / (fn + ".".join(flags) + ".out")
        cmd = ["g++", "-std=c++17", "-O1", *flags, "-I/usr/local/cuda/include", str(dst), "-o", str(exe),
               "-L" + os.path.join(ROOT, "parallel-gcn_b200"), "-lgcn_b200", "-L/usr/local/cuda/lib64", "-lcudart",
               "-Wl,-rpath," + os.path.join(ROOT, "parallel-gcn_b200")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]


def test_binary_container_roundtrip_is_bit_identical(tmp_path):
    """gcnb_dataset_save / gcnb_dataset_load: the stored dataset equals the parsed one byte for byte;
    truncated or foreign files are rejected."""
    import importlib
    import __graft_entry__ as ge
    ge.load_package()
    eng = importlib.import_module("parallel_gcn_b200.engine")
    for name in ("cora", "citeseer"):
        path = tmp_path / (name + ".gcnb")
        ds = eng.parse_dataset(ROOT, name, save_to=path)
        back = eng.load_dataset(path)
        assert back is not None
        for k in eng.HostDataset.FIELDS:
            a, b = getattr(ds, k), getattr(back, k)
            assert a.dtype == b.dtype and np.array_equal(a.view(np.uint32), b.view(np.uint32)), (name, k)
        assert (ds.input_dim, ds.output_dim, ds.split_counts) == (back.input_dim, back.output_dim, back.split_counts)
        raw = path.read_bytes()
        assert len(raw) % 64 == 0 and raw[:7] == b"GCNBDS1"
        (tmp_path / "cut.gcnb").write_bytes(raw[: len(raw) // 2])
        assert eng.load_dataset(tmp_path / "cut.gcnb") is None
    (tmp_path / "junk.gcnb").write_bytes(b"x" * 4096)
    assert eng.load_dataset(tmp_path / "junk.gcnb") is None
    assert eng.load_dataset(tmp_path / "missing.gcnb") is None


def test_rng_history_counts_global_elements_in_64_bits():
    """a row-partitioned model consumes GLOBAL element counts (8 M nodes x 602 features > 2^32); the Philox
    group count must not wrap, and distinct consumer sizes up to the table's capacity must fit (deep models)."""
    import ctypes as C
    import importlib
    import __graft_entry__ as ge
    ge.load_package()
    b = importlib.import_module("parallel_gcn_b200.binding")
    lib = b.lib
    lib.gcnb_rng_history_consume.argtypes = [C.c_uint64]
    lib.gcnb_rng_history_descriptor.argtypes = [C.c_void_p]
    assert lib.gcnb_rng_history_reset() == 0
    big = 8_000_000 * 602            # 4.8e9 elements > 2^32
    for n in (big, big, (1 << 32) + 5, 7):
        assert lib.gcnb_rng_history_consume(n) == 0
    r = b.RngT()
    assert lib.gcnb_rng_history_descriptor(C.byref(r)) == 0
    got = {int(r.hist_groups[i]): int(r.hist_count[i]) for i in range(r.n_hist)}
    assert got == {(big + 3) // 4: 2, ((1 << 32) + 5 + 3) // 4: 1, 2: 1}
    assert (big + 3) // 4 > (big % (1 << 32) + 3) // 4, "the old 32-bit count would have recorded a truncated prefix"
    # 40 distinct consumer sizes (a 20-layer model) fit the descriptor
    lib.gcnb_rng_history_reset()
    for k in range(40):
        lib.gcnb_rng_history_consume(1000 + 4 * k)
    assert lib.gcnb_rng_history_descriptor(C.byref(r)) == 0 and r.n_hist == 40 and b.MAX_RNG_HIST >= 40
    lib.gcnb_rng_history_reset()
