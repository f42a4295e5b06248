"""The reference's tuning sweep (test/tuning_accuracy.cpp:56-196) as a throughput workload: the 2-layer slice of its grid
(hidden {8,16,32,64} x dropout {0,.2,.4,.6}^2 x weight decay {5e-5,5e-4,5e-3}, early stopping 10, up to 1000 epochs), `reps`
seeds per combination, run through gcnb_sweep_run with 1 worker (= the reference's own loop order, one model at a time) and
with several.  Prints one JSON line per worker count.

  python scripts/bench_sweep.py [--dataset cora] [--reps 2] [--workers 1,4,8,16] [--max-epochs 1000]
"""
import argparse
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", default="cora")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--workers", default="1,4,8,16")
    ap.add_argument("--max-epochs", type=int, default=1000)
    ap.add_argument("--hidden", default="8,16,32,64")
    a = ap.parse_args()
    import __graft_entry__ as ge
    ge.load_package()
    import importlib
    eng = importlib.import_module("parallel_gcn_b200.engine")
    rnd = random.Random(5489)  # (the reference draws its seeds from a default-constructed mt19937)
    trials = []
    for hidden in [int(x) for x in a.hidden.split(",")]:
        for wd in (5e-5, 5e-4, 5e-3):
            for d1 in (0.0, 0.2, 0.4, 0.6):
                for d2 in (0.0, 0.2, 0.4, 0.6):
                    for _ in range(a.reps):
                        trials.append(dict(hidden_dims=(hidden,), dropouts=(d1, d2), epochs=a.max_epochs, early_stopping=10,
                                           learning_rate=0.01, weight_decay=wd, seed=rnd.randrange(1 << 31)))
    eng.sweep_run((ROOT, a.dataset), trials[:8], workers=2)  # warm-up: module load, first-launch costs
    base = None
    for w in [int(x) for x in a.workers.split(",")]:
        res, wall = eng.sweep_run((ROOT, a.dataset), trials, workers=w)
        epochs = sum(r["epochs_run"] for r in res)
        key = [(r["epochs_run"], r["last_val_accuracy"], r["last_val_loss"]) for r in res]
        if base is None:
            base = (key, wall)
        print(json.dumps({"workload": "tuning sweep, %s, %d trials (2 layers, early stopping 10, <= %d epochs)" % (a.dataset, len(trials), a.max_epochs),
                          "workers": w, "wall_s": round(wall, 3), "trials_per_s": round(len(trials) / wall, 1), "epochs_total": epochs,
                          "epochs_per_s": round(epochs / wall, 1), "ms_per_epoch_amortised": round(1000 * wall / epochs, 4),
                          "speedup_vs_first": round(base[1] / wall, 2), "identical_to_first": key == base[0],
                          "mean_val_acc": round(sum(r["last_val_accuracy"] for r in res) / len(res), 4)}), flush=True)


if __name__ == "__main__":
    main()
