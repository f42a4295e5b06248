"""GPU end-to-end parity of the GCN driver (through the engine C ABI, host buffers in) against the oracle.

Bars (north_star): integer results (wrong counts => accuracies, masks) bit-exact; losses, logits and weights within
1e-5 relative in fp32 after one epoch with identical randomness; over several epochs the fp32 tolerance widens with
the number of Adam steps (rounding-order differences are amplified by training), stated per test."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests.util import assert_close, to_dev, to_np

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32, u32, u8 = np.float32, np.uint32, np.uint8


@pytest.fixture(scope="module")
def eng(gcnb, dev):
    import importlib
    return importlib.import_module("parallel_gcn_b200.engine")


def test_philox_kernels_match_real_curand_on_device(O, gcnb, dev):
    """the stateless Philox of csrc/philox.cuh == curand_init(seed,i,0)+curand_uniform4 executed on this GPU."""
    import torch
    from tests.native import build as nb
    devlib, _ = nb.build()
    lib = C.CDLL(devlib)
    n_states, n_draws, seed, p = 1000, 3, 19990304, 0.5
    ref = np.zeros((n_draws, n_states, 4), f32)
    assert lib.curand_ref_device(C.c_uint(seed), n_states, n_draws, ref.ctypes.data_as(C.c_void_p)) == 0
    size = n_states * 4 - 1
    for t in range(n_draws):
        hist = [(size, t)] if t else []
        # dropout keep decisions: u >= p
        x = torch.ones(size, device=dev)
        m = torch.zeros(size, dtype=torch.uint8, device=dev)
        gcnb.dropout_fwd(x, m, p, rng=gcnb.make_rng(seed, hist))
        assert (to_np(m) == (ref[t].ravel()[:size] >= f32(p))).all(), t
        # glorot values: (u - 0.5) * scale in double
        w = torch.empty(size, device=dev)
        gcnb.glorot(w, 100, 28, gcnb.make_rng(seed, hist))
        scale = float(np.sqrt(f32(6.0) / f32(128))) * 2  # sqrtf in fp32, widened
        want = ((ref[t].ravel()[:size].astype(np.float64) - 0.5) * (float(f32(np.sqrt(f32(6.0 / 128)))) * 2)).astype(f32)
        assert (to_np(w).view(u32) == want.view(u32)).all(), t
        # and the oracle's device flavour agrees with the device
        u = np.zeros(4, f32)
        for i in (0, 1, 17, n_states - 1):
            O.lib.orc_curand_uniform4(seed, i, t, 1, O._p(u))
            assert (u.view(u32) == ref[t, i].view(u32)).all()


def _run_pair(O, eng, ds_o, ds_e, epochs, hidden=(16,), dropouts=(0.5, 0.5), reorder=True, seed=19990304, wd=5e-4):
    og = O.OracleGCN(ds_o, hidden_dims=hidden, dropouts=dropouts, flavour="ref_gpu", seed=seed, weight_decay=wd)
    g = eng.GCN(ds_e, hidden_dims=hidden, dropouts=dropouts, seed=seed, reorder=reorder, weight_decay=wd)
    for l in range(len(hidden) + 1):  # Glorot through Philox: bit-exact initial weights
        assert (g.weight(l).view(u32) == og.W[l].view(u32)).all(), "glorot layer %d" % l
    hist = []
    for ep in range(epochs):
        to, te = og.train_epoch(), g.train_epoch()
        vo, ve = og.eval(2), g.eval(2)
        hist.append((to, te, vo, ve))
    return og, g, hist


@pytest.mark.parametrize("name", ["cora", "citeseer"])
@pytest.mark.parametrize("reorder", [True, False])
def test_training_matches_oracle_ref_gpu_flavour(O, eng, datasets, name, reorder):
    """same seed => same Philox weights and dropout masks as the reference GPU code would draw; compare 10 epochs."""
    ds_o = datasets[name]
    ds_e = eng.parse_dataset(ROOT, name)
    og, g, hist = _run_pair(O, eng, ds_o, ds_e, 10, reorder=reorder)
    for ep, (to, te, vo, ve) in enumerate(hist):
        tol = 1e-5 * (1 + ep)  # widens with the number of optimizer steps
        assert abs(te[0] - to[0]) <= tol * abs(to[0]), ("train loss", ep, te, to)
        assert abs(ve[0] - vo[0]) <= tol * abs(vo[0]), ("val loss", ep, ve, vo)
        if ep < 3:  # integer results: identical predictions while weights still agree to ~1e-6
            assert te[1] == pytest.approx(to[1], abs=1e-7) and ve[1] == pytest.approx(vo[1], abs=1e-7), (ep, te, to, ve, vo)
        else:
            assert abs(te[1] - to[1]) < 5e-3 and abs(ve[1] - vo[1]) < 5e-3
    for l in range(2):
        assert_close(g.weight(l), og.W[l], rtol=2e-4, atol=2e-6, what="W%d after 10 epochs" % l)
    g.close()


@pytest.mark.parametrize("name", ["cora", "citeseer"])
def test_first_epoch_tensors(O, eng, datasets, name):
    """one training pass: logits, weight gradients and updated weights within 1e-5 relative (+ cancellation floor)."""
    ds_o, ds_e = datasets[name], eng.parse_dataset(ROOT, name)
    for reorder in (False, True):
        og = O.OracleGCN(ds_o, flavour="ref_gpu")
        g = eng.GCN(ds_e, reorder=reorder)
        to, te = og.train_epoch(), g.train_epoch()
        assert abs(te[0] - to[0]) <= 1e-5 * abs(to[0]) and te[1] == pytest.approx(to[1], abs=1e-7)
        assert_close(g.logits().ravel(), og.trace["logits"], rtol=1e-5, what="shifted logits")
        for l in range(2):
            assert_close(g.weight_grad(l), og.wgrads[l], rtol=1e-5, what="dW%d" % l)
            # Adam's first step is lr*g/(|g|+eps): for gradients of the order of eps (1e-8) the update is
            # ill-conditioned in g, so weights get an absolute floor of 1e-6 (2e-5 of their scale) on top of 1e-5 rel.
            assert_close(g.weight(l), og.W[l], rtol=1e-5, atol=1e-6, what="W%d" % l)
        # argmax predictions bit-exact
        assert (g.logits().argmax(1) == og.trace["logits"].reshape(-1, ds_o.output_dim).argmax(1)).all()
        g.close()


def test_training_matches_reference_cpu_with_injected_randomness(O, eng, datasets):
    """identical initial weights and dropout masks injected from the reference CPU code's xorshift stream (shared libc
    seed): the engine then tracks the reference's own published training curve (tests/golden/ref_cpu_training.json,
    generated from the unmodified hpdga-spring23 sources)."""
    import json
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_cpu_training.json")))["cora"]
    ds_o, ds_e = datasets["cora"], eng.parse_dataset(ROOT, "cora")
    N, H, fnnz = ds_o.num_nodes, 16, len(ds_o.f_indices)
    O.lib.orc_libc_srand(1)
    O.lib.orc_xorshift_seed_from_libc()
    w0, w1 = np.empty(ds_o.input_dim * H, f32), np.empty(H * ds_o.output_dim, f32)
    O.lib.orc_glorot_xorshift(w0.size, ds_o.input_dim, H, O._p(w0))
    O.lib.orc_glorot_xorshift(w1.size, H, ds_o.output_dim, O._p(w1))
    g = eng.GCN(ds_e)
    g.set_weight(0, w0); g.set_weight(1, w1)
    for ep in range(10):
        m_in, m_h = np.empty(fnnz, u8), np.empty(N * H, u8)
        O.lib.orc_dropout_mask_xorshift(fnnz, 0.5, O._p(m_in))   # draw order of hpdga-spring23/src/gcn.cpp:179-186
        O.lib.orc_dropout_mask_xorshift(N * H, 0.5, O._p(m_h))
        g.set_mask(0, m_in); g.set_mask(1, m_h)
        t, v = g.train_epoch(), g.eval(2)
        want = gold["epochs"][ep]
        tol = 2e-5 * (1 + ep)
        assert abs(t[0] - want[0]) <= tol * want[0] and abs(v[0] - want[2]) <= tol * want[2], (ep, t, v, want)
        assert abs(t[1] - want[1]) < 2e-3 and abs(v[1] - want[3]) < 4e-3, (ep, t, v, want)
    assert_close(g.weight(1), np.array(gold["w1_after_10_epochs"], f32), rtol=5e-4, atol=5e-6, what="W1 vs reference CPU")
    g.close()


def test_three_layer_model_and_zero_dropout(O, eng, datasets):
    """arbitrary L (src/gcn.cu:85-112), widths that exercise both associations, p = 0 still consuming RNG draws."""
    ds_o, ds_e = datasets["cora"], eng.parse_dataset(ROOT, "cora")
    for hidden, drops in (((32, 8), (0.5, 0.0, 0.3)), ((8, 24), (0.0, 0.2, 0.2))):
        og, g, hist = _run_pair(O, eng, ds_o, ds_e, 4, hidden=hidden, dropouts=drops, seed=1382895624, wd=5e-5)
        for ep, (to, te, vo, ve) in enumerate(hist):
            assert abs(te[0] - to[0]) <= 2e-5 * (1 + ep) * abs(to[0]), (hidden, ep, te, to)
            assert abs(ve[0] - vo[0]) <= 2e-5 * (1 + ep) * abs(vo[0]), (hidden, ep, ve, vo)
        g.close()


def test_run_output_format_and_early_stopping(eng, capfd):
    """GCN::run(): stdout lines byte-compatible with src/gcn.cu:371-372,432-433; early stopping rule of :377-394."""
    ds = eng.parse_dataset(ROOT, "cora")
    g = eng.GCN(ds, epochs=30, early_stopping=3, quiet=False, lr=0.2)
    res = g.run()
    out = capfd.readouterr().out
    lines = [l for l in out.splitlines() if l.startswith("epoch=")]
    pat = re.compile(r"^epoch=(\d+) train_loss=\d+\.\d{5} train_acc=\d\.\d{5} val_loss=(\d+\.\d{5}) val_acc=\d\.\d{5} time=\d+\.\d{5}$")
    assert lines and all(pat.match(l) for l in lines), lines[:3]
    assert re.search(r"^TMR_TRAIN average time: \d+\.\d{3}ms$", out, re.M)
    assert re.search(r"^test_loss=\d+\.\d{5} test_acc=\d\.\d{5} time=\d+\.\d{5}$", out, re.M) and "total time: " in out
    vals = [float(pat.match(l).group(2)) for l in lines]
    n = len(lines)
    assert res["epochs"] == n
    if n < 30:  # stopped early: last val loss above the mean of the last 3 (including itself)
        assert "Early stopping..." in out
        assert vals[-1] > sum(vals[-3:]) / 3 - 1e-5
    for k in range(3, n):  # ... and no earlier epoch satisfied the rule
        assert not (vals[k - 1] > sum(vals[k - 3:k]) / 3 + 1e-5) or k == n
    g.close()


def test_dense_feature_path_matches_sparse_path(O, eng):
    """an all-columns feature CSR (how Reddit parses) takes the dense kernels; results equal the CSR path's."""
    n, F, Cn = 700, 24, 5
    ds = eng.synth_dataset(n, 6000, F, Cn, n_blocks=4, seed=5)
    ods = O.Dataset(g_indptr=ds.g_indptr, g_indices=ds.g_indices, f_indptr=ds.f_indptr, f_indices=ds.f_indices,
                    f_value=ds.f_value, label=ds.label, split=ds.split, input_dim=F, output_dim=Cn)
    og, g, hist = _run_pair(O, eng, ods, ds, 3, hidden=(16,), dropouts=(0.5, 0.5))
    for ep, (to, te, vo, ve) in enumerate(hist):
        assert abs(te[0] - to[0]) <= 2e-5 * (1 + ep) * abs(to[0]) and abs(ve[0] - vo[0]) <= 2e-5 * (1 + ep) * abs(vo[0])
    assert_close(g.weight(0), og.W[0], rtol=1e-4, atol=1e-6, what="W0 dense path")
    g.close()


def test_dist_driver_single_rank_matches_engine(O, eng, gcnb, dev, datasets):
    """the multi-GPU driver (parallel_gcn_b200.dist.DistGCN, CUDA backend) at world size 1 reproduces the C++ engine:
    same kernels, same Philox bookkeeping.  (world > 1 choreography: tests/test_dist_cpu.py with gloo; N-GPU run:
    scripts/dist_check.py under torchrun.)"""
    import importlib
    dmod = importlib.import_module("parallel_gcn_b200.dist")
    for name in ("cora", "citeseer"):
        ds = eng.parse_dataset(ROOT, name)
        part = dmod.partition_dataset(ds, 0, 1)
        dg = dmod.DistGCN(part, dmod.CudaOps(gcnb, dev), dmod.Comm(None, 0, 1))
        g = eng.GCN(ds)
        for ep in range(4):
            a, b = dg.train_epoch(), g.train_epoch()
            va, vb = dg.eval(2), g.eval(2)
            assert abs(a[0] - b[0]) <= 1e-6 * abs(b[0]) and a[1] == pytest.approx(b[1], abs=1e-7), (name, ep, a, b)
            assert abs(va[0] - vb[0]) <= 1e-6 * abs(vb[0]) and va[1] == pytest.approx(vb[1], abs=1e-7)
        for l in range(2):
            assert_close(to_np(dg.W[l]), g.weight(l), rtol=1e-6, atol=1e-7, what="dist W%d" % l)
        g.close()


def test_native_partitioned_engine_single_rank_matches_engine(eng, datasets):
    """the row-partitioned native engine (gcnb_gcn_create_partitioned, csrc/comm.cu) on ONE rank: slab-padded buffers,
    global RNG offsets and the (degenerate) collectives must reproduce the plain engine bit for bit; the N-rank case is
    scripts/dist_check_native.py (needs N GPUs)."""
    import importlib
    dmod = importlib.import_module("parallel_gcn_b200.dist")
    ds = eng.parse_dataset(ROOT, "citeseer")
    comm = eng.Comm(0, 1)
    # one model at a time: the Philox consumption history is process-wide (Variable::rng_history, like the reference's
    # static state array), a second constructor restarts it
    g = eng.GCN(eng.PartDataset(dmod.partition_dataset(ds, 0, 1)), comm=comm)
    got = [(g.train_epoch(), g.eval(2)) for _ in range(3)]
    gw = [g.weight(l) for l in range(2)]
    g.close()
    comm.close()
    h = eng.GCN(ds)
    want = [(h.train_epoch(), h.eval(2)) for _ in range(3)]
    hw = [h.weight(l) for l in range(2)]
    h.close()
    assert got == want
    for a, b in zip(gw, hw):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name", ["cora", "synthetic"])
def test_renumbered_dataset_is_the_same_model(eng, datasets, name):
    """locality reordering (host/src/reorder.cpp): a dataset renumbered by gcnb_reorder_communities is the same learning
    problem -- with the same weights, evaluation loss / accuracy agree and the logits, mapped back with
    gcnb_unpermute_rows, are the original ones (fp32 summation order inside a row changes: 1e-5)."""
    if name == "cora":
        ds = eng.parse_dataset(ROOT, "cora")
    else:
        ds = eng.synth_dataset(30000, 30000 * 25, 32, 6, n_blocks=8, seed=11)
    perm, n_comm = eng.reorder_communities(ds.g_indptr, ds.g_indices)
    assert n_comm >= 1 and np.array_equal(np.sort(perm), np.arange(ds.num_nodes, dtype=np.uint32))
    pds = eng.permute_dataset(ds, perm)
    g = eng.GCN(ds)
    w = [g.weight(l) for l in range(2)]
    ev = [g.eval(s) for s in (1, 2, 3)]
    logits = g.logits()
    g.close()
    h = eng.GCN(pds)
    for l in range(2):
        h.set_weight(l, w[l])
    evp = [h.eval(s) for s in (1, 2, 3)]
    plogits = h.logits()
    h.close()
    for a, b in zip(ev, evp):
        assert abs(a[0] - b[0]) <= 1e-5 * abs(a[0]) and abs(a[1] - b[1]) <= 1e-6
    assert_close(eng.permute_rows(plogits, perm, inverse=True), logits, rtol=1e-5, what="un-permuted logits")


@pytest.mark.parametrize("name", ["cora", "citeseer", "dense"])
def test_cuda_graph_replay_is_bit_identical_to_eager(eng, datasets, name):
    """small datasets replay captured epochs (CUDA graphs) with the per-epoch kernel arguments -- Philox descriptors,
    Adam step size -- patched into the instantiated graph: every loss, accuracy and weight equals the eager run's bits."""
    def make():
        if name == "dense":
            return eng.synth_dataset(3000, 40000, 24, 5, n_blocks=4, seed=9)
        return eng.parse_dataset(ROOT, name)

    def run(use_graph):
        g = eng.GCN(make(), hidden_dims=(16,), dropouts=(0.5, 0.5))
        g.set_cuda_graph(use_graph)
        assert g.uses_cuda_graph() == use_graph
        hist = []
        for _ in range(6):
            hist.append((g.train_epoch(), g.eval(2)))
        hist.append((g.eval(3), g.eval(1)))
        w = [g.weight(l) for l in range(2)]
        launches = g.launches_per_epoch()
        g.close()
        return hist, w, launches

    eager, graph = run(False), run(True)
    assert eager[0] == graph[0]
    for a, b in zip(eager[1], graph[1]):
        assert np.array_equal(a, b)
    assert eager[2] == graph[2]


def test_tuning_sweep_is_independent_of_concurrency(eng):
    """gcnb_sweep_run (the reference's test/tuning_accuracy.cpp loop as a throughput workload): trials running side by side
    on several host threads and streams give bit for bit the results of running them one after the other, and those of the
    plain `CudaParams::SEED = seed; GCN gcn{...}; gcn.run()` a caller of the reference would write."""
    trials = []
    seeds = [1804289383, 846930886, 1681692777, 1714636915]
    for hidden in ((8,), (16,), (32, 32)):
        for d1, d2 in ((0.0, 0.2), (0.6, 0.4)):
            for seed in seeds[:2]:
                trials.append(dict(hidden_dims=hidden, dropouts=(d1,) + (d2,) * len(hidden), epochs=40, early_stopping=10,
                                   learning_rate=0.01, weight_decay=5e-4, seed=seed))
    trials.append(dict(hidden_dims=(16,), dropouts=(0.5, 0.5), epochs=25, early_stopping=0, seed=seeds[2]))  # pipelined replays
    one, _ = eng.sweep_run((ROOT, "cora"), trials, workers=1)
    many, _ = eng.sweep_run((ROOT, "cora"), trials, workers=6)
    keys = ("last_val_accuracy", "last_val_loss", "last_train_loss", "epochs_run")
    for t, a, b in zip(trials, one, many):
        assert all(a[k] == b[k] for k in keys), (t, a, b)
        assert 1 <= a["epochs_run"] <= t["epochs"] and 0.0 <= a["last_val_accuracy"] <= 1.0 and np.isfinite(a["last_val_loss"])
    assert len({r["last_val_accuracy"] for r in one}) > 3  # different seeds / models really are different runs
    ds = eng.parse_dataset(ROOT, "cora")
    for i in (0, 5, len(trials) - 1):
        t = trials[i]
        g = eng.GCN(ds, hidden_dims=t["hidden_dims"], dropouts=t["dropouts"], epochs=t["epochs"], early_stopping=t["early_stopping"],
                    lr=t.get("learning_rate", 0.01), weight_decay=t.get("weight_decay", 5e-4), seed=t["seed"], quiet=True)
        res = g.run()
        g.close()
        assert res["epochs"] == one[i]["epochs_run"] and res["last_val_acc"] == one[i]["last_val_accuracy"], (t, res, one[i])
