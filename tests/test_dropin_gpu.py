"""The C++ drop-in EXECUTED on the GPU:

  * the product's module classes (host/src/module.cpp: Dropout, SparseMatmul, GraphSum, ReLU, Matmul, CrossEntropyLoss)
    composed by hand exactly as the reference's GCN constructor composes its own (src/gcn.cu:47-142), one training
    forward + backward and one evaluation forward, every tensor against the oracle's module outputs
    (tests/native/module_chain.cpp, built on the spot with g++);
  * the reference's OWN src/main.cpp, compiled unchanged against the product's headers and linked with libgcn_b200.so
    (oracle/_ref/dropin_main[_part2], built where /root/reference exists), run on cora: its per-epoch stdout must be the
    engine's training curve, and for the Part-2 binary the parameter file, early stopping and the test line as well."""
import os
import re
import subprocess

import numpy as np
import pytest

from tests.util import assert_close

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32, u32 = np.float32, np.uint32


@pytest.fixture(scope="module")
def eng(gcnb, dev):
    import importlib
    return importlib.import_module("parallel_gcn_b200.engine")


@pytest.mark.parametrize("name", ["cora", "citeseer"])
def test_module_chain_assembled_by_hand_matches_oracle(O, dev, datasets, tmp_path, name):
    from tests.native import build as nb
    exe = nb.build_module_chain()
    r = subprocess.run([exe, name, str(tmp_path)], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "module_chain ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    meta = open(tmp_path / "meta.txt").read().split()
    N, F, H, Cn, n_train, n_val = (int(x) for x in meta[:6])
    loss_tr_sum, loss_va_sum = float.fromhex(meta[6]), float.fromhex(meta[7])
    ds = datasets[name]
    assert (N, F, Cn) == (ds.num_nodes, ds.input_dim, ds.output_dim) and (n_train, n_val) == ds.split_counts()[:2]

    def load(fn):
        return np.fromfile(tmp_path / fn, f32)

    og = O.OracleGCN(ds, flavour="ref_gpu")  # same default seed as CudaParams::SEED => same Philox weights and masks
    assert (load("w0_init.f32").view(u32) == og.W[0].view(u32)).all() and (load("w1_init.f32").view(u32) == og.W[1].view(u32)).all()
    ev = O.OracleGCN(ds, flavour="ref_gpu")
    val_loss, _ = ev.forward(2, False)       # evaluation with the initial weights (the chain takes no optimizer step)
    assert_close(load("val_logits.f32"), ev.trace["logits"], what="module chain: evaluation logits")
    l2 = float(np.sum(og.W[0].astype(np.float64) ** 2))
    assert abs(loss_va_sum / n_val + 5e-4 * l2 / 2 - val_loss) <= 1e-5 * abs(val_loss)
    train_loss, _ = og.forward(1, True)
    assert abs(loss_tr_sum / n_train + 5e-4 * l2 / 2 - train_loss) <= 1e-5 * abs(train_loss)
    assert_close(load("train_logits.f32"), og.trace["logits"], what="module chain: training logits (shifted in place)")
    assert_close(load("train_hidden.f32"), og.trace["var2_0"], what="module chain: hidden activations after ReLU + dropout")
    assert_close(load("dlogits.f32"), og.trace["grad"], what="module chain: d loss / d logits")
    og.backward_and_step()
    assert_close(load("dw1.f32"), og.wgrads[1], what="module chain: dW1")
    assert_close(load("dw0.f32"), og.wgrads[0], what="module chain: dW0")


EPOCH = re.compile(r"epoch=(\d+) train_loss=([\d.]+) train_acc=([\d.]+) val_loss=([\d.]+) val_acc=([\d.]+)")


def _run_main(exe, tmp_path, param_text):
    if not os.path.exists(exe):
        pytest.skip("%s was not built (make -C oracle needs /root/reference and libgcn_b200.so)" % os.path.relpath(exe, ROOT))
    os.symlink(os.path.join(ROOT, "data"), tmp_path / "data")
    (tmp_path / "parameters").mkdir()
    (tmp_path / "parameters" / "parameters_cora.txt").write_text(param_text)
    r = subprocess.run([exe, "cora"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = [tuple(float(x) for x in m.groups()) for m in EPOCH.finditer(r.stdout)]
    assert rows and [int(x[0]) for x in rows] == list(range(1, len(rows) + 1)), r.stdout[-2000:]
    return r.stdout, rows


PARAMS = """n_layers = 2
hidden_dims = 72
dropouts = 0.4,0.2
epochs = 40
early_stopping = 10
learning_rate = 0.01
weight_decay = 5e-5
beta1 = 0.9
beta2 = 0.999
eps = 1e-8
num_blocks_factor = 2
num_threads = 1024
seed = 1382895624
"""


def _same_curve(rows, g, what):
    for ep, row in enumerate(rows):
        tl, vl = g.train_epoch(), g.eval(2)
        want = (tl[0], tl[1], vl[0], vl[1])
        for a, b in zip(row[1:], want):  # printed with 5 decimals; same library, same seed => the same numbers
            assert abs(a - b) <= 1.5e-5, (what, ep + 1, row, want)


def test_reference_main_part1_runs_against_the_library(eng, tmp_path):
    """Part-1 binary: the parameter file's model keys are ignored: hidden 16, dropout 0.5, 100 epochs"""
    out, rows = _run_main(os.path.join(ROOT, "oracle", "_ref", "dropin_main"), tmp_path, PARAMS)
    assert len(rows) == 100 and "test_loss=" in out and "total time" in out
    g = eng.GCN(eng.parse_dataset(ROOT, "cora"))
    _same_curve(rows, g, "part 1")
    te = g.eval(3)
    m = re.search(r"test_loss=([\d.]+) test_acc=([\d.]+)", out)
    assert abs(float(m.group(1)) - te[0]) <= 1.5e-5 and abs(float(m.group(2)) - te[1]) <= 1.5e-5
    g.close()


def test_reference_main_part2_reads_the_parameter_file_and_stops_early(eng, tmp_path):
    out, rows = _run_main(os.path.join(ROOT, "oracle", "_ref", "dropin_main_part2"), tmp_path, PARAMS)
    assert "hidden_dims: 72" in out and len(rows) <= 40
    g = eng.GCN(eng.parse_dataset(ROOT, "cora"), hidden_dims=(72,), dropouts=(0.4, 0.2), epochs=40, early_stopping=10, lr=0.01,
                weight_decay=5e-5, seed=1382895624)
    _same_curve(rows, g, "part 2")
    g.close()
    if len(rows) < 40:  # the rule of src/gcn.cu:377-394 fired: last val loss above the mean of the last 10
        assert "Early stopping" in out
        vals = [r[3] for r in rows]
        assert vals[-1] > sum(vals[-10:]) / 10 - 1e-5
