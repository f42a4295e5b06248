"""Bit-tile GraphSum probe (csrc/spmm_bittile.cu): correctness against a float64 product on hand-made dense patterns, then
correctness against the generic kernel + timing on community graphs.  Every step appends one JSON line to --out and
flushes, so a hang in a later step loses nothing.  Run under `timeout`.

  python scripts/probe_bittile.py --stage small           # layout / pipeline diagnostics, seconds
  python scripts/probe_bittile.py --stage graph --scale 8 # 1/8 Reddit-shape community graph vs the generic kernel
  python scripts/probe_bittile.py --stage graph --scale 1 # the bench graph
  ... --chunk 128 | --rb 2 | GCNB_BT_UNIFIED=1            # second-generation MMA kernel (128-column tiles / 256-row items /
                                                          # 64-column tiles with unified stage barriers)
Round-2 checklist: GCNB_TEST_BITTILE_WIDE=1 GCNB_TEST_BITTILE_ENGINE=1 pytest tests/test_zz_bittile_gpu.py -m gpu; this
probe for the three shapes at --scale 1; then `ncu --set full --import-source on -k regex:bt_mma -c 1` on the best one
(warp-state stalls per role: expanders = warps 0-7, epilogue 8-11, producer 12, MMA issuer 13).
"""
import argparse
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import __graft_entry__ as ge

ge.load_package()
from parallel_gcn_b200 import binding as gcnb
from tests.test_bittile_cpu import pack_b

ap = argparse.ArgumentParser()
ap.add_argument("--stage", default="small")
ap.add_argument("--scale", type=int, default=8)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--min-tile-nnz", type=int, default=0)
ap.add_argument("--staged", type=int, default=1, help="also time the window-staged path")
ap.add_argument("--chunk", type=int, default=0, help="columns per tile: 64 (default) or 128; GCNB_BT_UNIFIED=1 selects the\n"
                "second-generation kernel for 64")
ap.add_argument("--rb", type=int, default=0, help="128-row blocks per item: 1 (default) or 2 (items of 256 rows share a B' stage)")
ap.add_argument("--out", default="gpurun_out/probe_bittile.jsonl")
args = ap.parse_args()
dev = torch.device("cuda:0")
gcnb.device_check()
log = open(args.out, "a")


def emit(**kw):
    log.write(json.dumps(kw) + "\n")
    log.flush()
    print(json.dumps(kw), flush=True)


def dense_case(name, n, density, seed, cols_used=None, dump=False):
    rng = np.random.default_rng(seed)
    M = rng.random((n, n)) < density
    if cols_used is not None:
        M[:, cols_used:] = False
    rs = (0.5 + rng.random(n)).astype(np.float32)
    cs = (0.5 + rng.random(n)).astype(np.float32)
    rows, cols = np.nonzero(M)
    indptr = np.zeros(n + 1, np.uint32)
    indptr[1:] = np.cumsum(M.sum(1))
    indices = cols.astype(np.uint32)
    values = (rs[rows] * cs[cols]).astype(np.float32)
    B = rng.standard_normal((n, 16)).astype(np.float32)
    plan = gcnb.BitTilePlan(indptr, indices, values, n, rs, cs, min_tile_nnz=1, chunk_cols=args.chunk, row_blocks=args.rb)
    info = plan.info()
    Bd = torch.from_numpy(B).to(dev)
    packed = plan.debug_pack(Bd)
    want_packed, _ = pack_b(B, cs)
    pack_ok = bool(np.array_equal(packed, want_packed))
    emit(case=name, step="pack", pack_ok=pack_ok, info=info)
    out = torch.full((n, 16), float("nan"), device=dev)
    plan.spmm16(Bd, out)
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)
    ref = rs[:, None].astype(np.float64) * (M.astype(np.float64) @ (cs[:, None].astype(np.float64) * B))
    err = np.abs(got - ref)
    scale = np.abs(ref).max()
    emit(case=name, step="spmm", max_err=float(np.nanmax(err)), nan=int(np.isnan(got).sum()), scale=float(scale),
         ok=bool(np.nanmax(err) <= 2e-6 * scale and not np.isnan(got).any()))
    if dump and not (np.nanmax(err) <= 2e-6 * scale):
        np.savez("gpurun_out/probe_bittile_%s.npz" % name, got=got, ref=ref, M=M, B=B, rs=rs, cs=cs, packed=packed)
    plan.close()


def graph_case(scale):
    from parallel_gcn_b200 import engine as eng
    n, m = 232965 // scale, 57307946 // scale
    t0 = time.time()
    indptr, indices = eng.synth_graph(n, m, n_blocks=max(2, 50 // scale))
    values = eng.synth_graph_values(indptr, indices, 0, np.diff(indptr).astype(np.uint32))  # parser.cpp:164-181 formula
    t1 = time.time()
    plan = gcnb.BitTilePlan(indptr, indices, values, n, min_tile_nnz=args.min_tile_nnz, chunk_cols=args.chunk,
                            row_blocks=args.rb)
    t2 = time.time()
    info = plan.info()
    emit(case="graph/%d" % scale, step="plan", n=n, nnz=int(indices.size), gen_s=t1 - t0, plan_s=t2 - t1, info=info)
    d_indptr, d_indices = torch.from_numpy(indptr.astype(np.int32)).to(dev), torch.from_numpy(indices.astype(np.int32)).to(dev)
    d_values = torch.from_numpy(values).to(dev)
    B = torch.randn(n, 16, device=dev)
    ref_plan = gcnb.SpmmPlan(d_indptr, d_indices, n)
    ref = torch.empty(n, 16, device=dev)
    ref_plan.spmm(d_values, B, ref, 16)
    out = torch.full((n, 16), float("nan"), device=dev)
    plan.spmm16(B, out)
    torch.cuda.synchronize()
    err = (out - ref).abs()
    tol = 1e-5 * ref.abs() + 1e-6 * ref.abs().max()
    emit(case="graph/%d" % scale, step="spmm", max_err=float(err.max()), scale=float(ref.abs().max()),
         bad=int((err > tol).sum()), nan=int(torch.isnan(out).sum()))

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters * 1e3

    us_bt = timeit(lambda: plan.spmm16(B, out))
    us_gen = timeit(lambda: ref_plan.spmm(d_values, B, ref, 16))
    alg = 4 * (n + 1) + 8 * indices.size + 8 * n * 16
    emit(case="graph/%d" % scale, step="time", bittile_us=us_bt, generic_us=us_gen, algorithmic_bytes=alg,
         bittile_gbs=alg / us_bt / 1e3, generic_gbs=alg / us_gen / 1e3)
    parts = {}
    for name, mask in (("pack", 1), ("mma", 2), ("remainder", 4), ("add", 8), ("pack+mma", 3), ("mma+remainder", 6), ("all", 15)):
        plan.debug_parts(mask)
        parts[name] = timeit(lambda: plan.spmm16(B, out))
    plan.debug_parts(15)
    emit(case="graph/%d" % scale, step="parts_us", **parts)
    if args.staged:
        ref_plan.stage(d_values, 16, indptr, indices)
        us_st = timeit(lambda: ref_plan.spmm(d_values, B, ref, 16))
        emit(case="graph/%d" % scale, step="time_staged", staged_us=us_st, staged_gbs=alg / us_st / 1e3,
             stage_info=ref_plan.stage_info())


if args.stage == "small":
    dense_case("one_tile", 128, 0.3, 1, cols_used=64, dump=True)
    dense_case("four_tiles", 256, 0.3, 2, dump=True)
    dense_case("chains", 1024, 0.2, 3)
    dense_case("ragged", 1000, 0.2, 4)
else:
    graph_case(args.scale)
