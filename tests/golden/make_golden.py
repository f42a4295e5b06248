"""Generates the committed golden fixtures from the REFERENCE's own CPU implementation (oracle/_ref, built in place
from /root/reference/hpdga-spring23 by oracle/Makefile).  Run in the build container:  python tests/golden/make_golden.py
The fixtures travel to the GPU box, where /root/reference does not exist."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

assert O.ref is not None, "oracle/_ref/libref_cpu.so missing: run `make -C oracle` with /root/reference present"
HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32

datasets, training = {}, {}
for name in ("cora", "citeseer"):
    h, ds = O.ref_parse_dataset(ROOT, name)
    ods = O.parse_dataset(os.path.join(ROOT, "data", name))
    gv = ods.graph_values()
    datasets[name] = dict(
        num_nodes=ds.num_nodes, graph_nnz=len(ds.g_indices), feat_nnz=len(ds.f_indices), input_dim=ds.input_dim,
        output_dim=ds.output_dim, split_counts=list(ods.split_counts()),
        unlabelled_train_nodes=[int(i) for i in np.nonzero((ds.split == 1) & (ds.label < 0))[0]],
        max_degree=int(np.diff(ds.g_indptr).max()), graph_value_sum=float(gv.astype(np.float64).sum()),
        fnv={k: O.fnv(getattr(ds, k)) for k in O.Dataset.FIELDS} | {"graph_value": O.fnv(gv)})
    # training curve of the unmodified reference CPU code, first GCN of a process (libc rand() seeded with 1)
    O.ref.ref_srand(1)
    g = O.ref.ref_gcn_create(h, 16, 0.5, 0.01, 5e-4, 100, 0)
    out = np.zeros(2, f32)
    epochs = []
    for ep in range(10):
        O.ref.ref_gcn_train_epoch(g, O._p(out)); t = [float(out[0]), float(out[1])]
        O.ref.ref_gcn_eval(g, 2, O._p(out)); epochs.append(t + [float(out[0]), float(out[1])])
    w0 = np.empty(O.ref.ref_gcn_variable_size(g, 2, 0), f32); O.ref.ref_gcn_variable_get(g, 2, 0, O._p(w0))
    w1 = np.empty(O.ref.ref_gcn_variable_size(g, 5, 0), f32); O.ref.ref_gcn_variable_get(g, 5, 0, O._p(w1))
    training[name] = dict(epochs=epochs, w0_fnv=O.fnv(w0), w1_fnv=O.fnv(w1), w1_after_10_epochs=[float(x) for x in w1])
    O.ref.ref_gcn_free(g)
    # the full 100-epoch run the reference's README/report quotes
    O.ref.ref_srand(1)
    g = O.ref.ref_gcn_create(h, 16, 0.5, 0.01, 5e-4, 100, 0)
    curve = []
    for ep in range(100):
        O.ref.ref_gcn_train_epoch(g, O._p(out)); t = [float(out[0]), float(out[1])]
        O.ref.ref_gcn_eval(g, 2, O._p(out)); curve.append(t + [float(out[0]), float(out[1])])
    O.ref.ref_gcn_eval(g, 3, O._p(out))
    training[name]["epoch_100"] = curve[-1]
    training[name]["epoch_50"] = curve[49]
    training[name]["test"] = [float(out[0]), float(out[1])]
    O.ref.ref_gcn_free(g)
    O.ref.ref_dataset_free(h)

json.dump(datasets, open(os.path.join(HERE, "datasets.json"), "w"), indent=1)
json.dump(training, open(os.path.join(HERE, "ref_cpu_training.json"), "w"), indent=1)
print("wrote", HERE, {k: (v["epoch_100"], v["test"]) for k, v in training.items()})
