"""Same-box A/B: the reference's OWN CUDA implementation (oracle/_ref/ref_gpu_bench = /root/reference/src/*.cu compiled for
sm_100 behind oracle/ref_gpu_harness.cu) against this engine, on the same dataset, same B200, same metric
(GCN::run()'s avg_epoch_time = train epoch + validation forward, ms).  Measurement infrastructure, not product code.

  python scripts/bench_ref_gpu.py --dataset cora --epochs 100 --reps 5
  python scripts/bench_ref_gpu.py --dataset reddit_shape --epochs 20 --reps 1       # BASELINE.json configs[2]
  python scripts/bench_ref_gpu.py --dataset reddit_shape --scale 8 ...               # 1/8-size debug run

The reference runs in its own process (its class names are the product's; a fault there must not take this one down)
on raw arrays dumped to a scratch directory -- its text parser would need hours for 115 M entries.  One JSON line per arm
and one with the ratio are appended to --out."""
import argparse
import importlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_gpu_bench")


def dump(ds, d):
    importlib.import_module("bench").dump_dataset_raw(ds, d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", default="cora", help="cora | citeseer | reddit_shape")
    ap.add_argument("--scale", type=int, default=1)
    ap.add_argument("--epochs", type=int, default=100)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--timeout", type=int, default=600, help="seconds for the reference process")
    ap.add_argument("--skip-ours", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_gpu_ab.jsonl"))
    args = ap.parse_args()
    ge.load_package()
    eng = importlib.import_module("parallel_gcn_b200.engine")
    if args.dataset == "reddit_shape":
        bench = importlib.import_module("bench")
        ds, w, gen_s = bench.make_dataset(eng, args.scale, pinned=False)
        label = "reddit_shape_synthetic/%d n=%d nnz=%d f=%d c=%d" % (args.scale, ds.num_nodes, len(ds.g_indices), w["f"], w["c"])
    else:
        ds = eng.parse_dataset(ROOT, args.dataset)
        label = args.dataset
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    out = open(args.out, "a")

    def emit(d):
        out.write(json.dumps(d) + "\n")
        out.flush()
        print(json.dumps(d), flush=True)

    ref = None
    if not os.path.exists(REF_BIN):
        emit({"impl": "reference_gpu", "unavailable": "oracle/_ref/ref_gpu_bench is not built (make -C oracle, needs /root/reference)"})
    else:
        scratch = tempfile.mkdtemp(prefix="gcnb_ref_gpu_")
        try:
            t0 = time.time()
            dump(ds, scratch)
            t_dump = time.time() - t0
            p = subprocess.run([REF_BIN, scratch, str(args.epochs), str(args.reps)], capture_output=True, text=True,
                               timeout=args.timeout)
            line = [l for l in p.stdout.splitlines() if l.startswith("{")]
            if line and json.loads(line[-1]).get("cuda_error") == "no error":
                ref = json.loads(line[-1])
                ref.update(dataset=label, dump_s=round(t_dump, 2))
                emit(ref)
            else:
                emit({"impl": "reference_gpu", "dataset": label, "failed": p.returncode, "stderr": p.stderr[-400:], "stdout": p.stdout[-400:]})
        except subprocess.TimeoutExpired:
            emit({"impl": "reference_gpu", "dataset": label, "failed": "timeout after %d s" % args.timeout})
        finally:
            shutil.rmtree(scratch, ignore_errors=True)
    if args.skip_ours:
        return
    import torch
    best = None
    for _ in range(args.reps):
        g = eng.GCN(ds, hidden_dims=(16,), dropouts=(0.5, 0.5), epochs=args.epochs)
        res = g.run()  # GCN::run(): avg_epoch_time (ms) of train epoch + validation forward, the reference's own timers' semantics
        torch.cuda.synchronize()
        g.close()
        if best is None or res["avg_epoch_ms"] < best["avg_epoch_ms"]:
            best = res
    ours = {"impl": "ours", "dataset": label, "epochs": args.epochs, "reps": args.reps, "best_avg_epoch_ms": best["avg_epoch_ms"],
            "total_s": best["total_s"], "last_val_accuracy": best["last_val_acc"]}
    emit(ours)
    if ref and ref.get("best_avg_epoch_ms"):
        emit({"dataset": label, "reference_gpu_ms_per_epoch": ref["best_avg_epoch_ms"], "ours_ms_per_epoch": ours["best_avg_epoch_ms"],
              "speedup_vs_reference_gpu_same_box": ref["best_avg_epoch_ms"] / ours["best_avg_epoch_ms"]})


if __name__ == "__main__":
    main()
