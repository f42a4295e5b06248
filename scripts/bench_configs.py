"""The other BASELINE.json configs on one GPU (they are parity-test cases, not the bench line): shipped datasets with the
reference's Part-1 defaults (configs[0], [1]) and the Reddit-shape graph with parameters/parameters_reddit.txt's model
(hidden 600, dropouts 0.0/0.1, wd 5e-5: configs[3]).  Prints one JSON line per config."""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.load_package()
eng = importlib.import_module("parallel_gcn_b200.engine")
import bench  # noqa: E402

from tests.util import pubmed_root  # noqa: E402


which = sys.argv[1:] or ["cora", "citeseer", "pubmed", "reddit600"]
for name in which:
    if name in ("cora", "citeseer", "pubmed"):
        ds = eng.parse_dataset(pubmed_root(ROOT) if name == "pubmed" else ROOT, name)
        g = eng.GCN(ds, epochs=100)
        g.train_epoch(); g.eval(2)
        r = g.timed_epochs(100, with_eval=True)
        out = dict(config=name + (" (shipped graph/split, synthetic svmlight features)" if name == "pubmed" else ""),
                   model="L2 H16 dropout .5/.5", ms_per_epoch=r["ms"] / 100, launches_per_step=r["launches"] / 100)
        g.close()
        g = eng.GCN(ds, epochs=100)
        t0 = time.perf_counter()
        rr = g.run()
        out.update(run_avg_epoch_ms_reference_style=rr["avg_epoch_ms"], run_wall_s=time.perf_counter() - t0,
                   last_val_acc=rr["last_val_acc"])
        g.close()
    else:
        ds, w, gen_s = bench.make_dataset(eng, 1, pinned=True)
        g = eng.GCN(ds, hidden_dims=(600,), dropouts=(0.0, 0.1), lr=0.01, weight_decay=5e-5, seed=w["seed"])
        g.finish_setup()  # attach the background-built GraphSum representation before timing
        for _ in range(2):
            g.train_epoch(); g.eval(2)
        r = g.timed_epochs(5, with_eval=True, time_graphsum=True)
        out = dict(config="reddit_shape H600 (parameters_reddit.txt model)", ms_per_epoch=r["ms"] / 5,
                   graphsum_ms_mean=r["graphsum_ms"] / max(1, r["graphsum_calls"]), graphsum_calls_per_step=r["graphsum_calls"] / 5,
                   launches_per_step=r["launches"] / 5, train=g.train_epoch(), val=g.eval(2))
        g.close()
    print(json.dumps(out), flush=True)
