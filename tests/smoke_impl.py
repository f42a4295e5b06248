"""__graft_entry__.smoke(): one small hot-path invocation on cuda:0 checked against the oracle."""
import importlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run():
    import torch
    import __graft_entry__ as ge
    ge.load_package()
    gcnb = importlib.import_module("parallel_gcn_b200.binding")
    eng = importlib.import_module("parallel_gcn_b200.engine")
    from oracle import oracle as O
    from tests.util import assert_close, to_dev, to_np

    assert torch.cuda.is_available(), "smoke() needs cuda:0"
    gcnb.device_check()
    dev = torch.device("cuda:0")
    ds = O.parse_dataset(os.path.join(ROOT, "data", "cora"))
    # 1) the north-star kernel: GraphSum on cora, d = 16
    n, dim = ds.num_nodes, 16
    x = np.random.default_rng(0).standard_normal((n, dim)).astype(np.float32)
    vals = ds.graph_values()
    want = np.empty((n, dim), np.float32)
    O.lib.orc_graphsum(n, dim, O._p(ds.g_indptr), O._p(ds.g_indices), O._p(vals), O._p(x), O._p(want))
    plan = gcnb.SpmmPlan(to_dev(ds.g_indptr, dev), to_dev(ds.g_indices, dev), n)
    out = torch.empty((n, dim), device=dev)
    plan.spmm(to_dev(vals, dev), to_dev(x, dev), out, dim)
    torch.cuda.synchronize()
    assert_close(to_np(out), want, what="smoke graphsum")
    # 2) one full training epoch + validation forward through the engine C ABI (host buffers in)
    og = O.OracleGCN(ds, flavour="ref_gpu")
    g = eng.GCN(eng.parse_dataset(ROOT, "cora"))
    (lo, ao), (le, ae) = og.train_epoch(), g.train_epoch()
    vo, ve = og.eval(2), g.eval(2)
    assert abs(le - lo) <= 1e-5 * abs(lo) and abs(ae - ao) < 1e-6, (le, lo, ae, ao)
    assert abs(ve[0] - vo[0]) <= 1e-5 * abs(vo[0]) and abs(ve[1] - vo[1]) < 1e-6, (ve, vo)
    print("smoke ok: graphsum parity, epoch loss %.6f (oracle %.6f), val acc %.4f, %d launches/epoch" %
          (le, lo, ve[1], g.launches_per_epoch()))
    g.close()
    # 3) GraphSum on a community graph through the tcgen05 bit-tile path (csrc/spmm_bittile.cu)
    from tests.test_bittile_cpu import gcn_graph
    ip, ix, gv = gcn_graph(np.random.default_rng(1), 1500, 3, 40, 3)
    xb = np.random.default_rng(2).standard_normal((1500, 16)).astype(np.float32)
    want_b = np.empty((1500, 16), np.float32)
    O.lib.orc_graphsum(1500, 16, O._p(ip), O._p(ix), O._p(gv), O._p(xb), O._p(want_b))
    bt = gcnb.BitTilePlan(ip, ix, gv, 1500, min_tile_nnz=64)
    out_b = torch.empty((1500, 16), device=dev)
    bt.spmm16(to_dev(xb, dev), out_b)
    torch.cuda.synchronize()
    assert bt.info()["n_tiles"] > 0
    assert_close(to_np(out_b), want_b, what="smoke bit-tile graphsum")
    bt.close()
    print("smoke ok: bit-tile GraphSum (tcgen05) parity on a community graph")
