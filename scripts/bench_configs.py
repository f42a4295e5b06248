"""The other BASELINE.json configs on one GPU (they are parity-test cases, not the bench line): shipped datasets with the
reference's Part-1 defaults (configs[0], [1]) and the Reddit-shape graph with parameters/parameters_reddit.txt's model
(hidden 600, dropouts 0.0/0.1, wd 5e-5: configs[3]).  Prints one JSON line per config (the same records bench.py attaches
to its line as `other_configs`)."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.load_package()
eng = importlib.import_module("parallel_gcn_b200.engine")
import bench  # noqa: E402

which = sys.argv[1:] or ["cora", "citeseer", "pubmed", "reddit600"]
for name in which:
    if name in ("cora", "citeseer", "pubmed"):
        out = bench.small_config_record(eng, name)
    else:
        ds, w, gen_s = bench.make_dataset(eng, 1, pinned=True)
        out = bench.wide_config_record(eng, ds, w)
    print(json.dumps(out), flush=True)
