"""torchrun --nproc-per-node N scripts/dist_check_native.py : N-GPU parity of the NATIVE row-partitioned engine
(host/src/gcn.cpp + csrc/comm.cu, NCCL collectives issued by the engine) against the oracle on cora / citeseer, and on
a community graph large enough for the window-staged GraphSum.  torch.distributed only ships the NCCL unique id."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.load_package()
eng = importlib.import_module("parallel_gcn_b200.engine")
dmod = importlib.import_module("parallel_gcn_b200.dist")
from oracle import oracle as O  # noqa: E402  (checker only)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
comm = dmod.make_comm(eng, dist, rank, world, dev)
ok = True
for name in ("cora", "citeseer"):
    ds = eng.parse_dataset(ROOT, name)
    g = eng.GCN(eng.PartDataset(dmod.partition_dataset(ds, rank, world)), comm=comm)
    og = O.OracleGCN(O.parse_dataset(os.path.join(ROOT, "data", name)), flavour="ref_gpu") if rank == 0 else None
    if rank == 0:
        print(name, "halo", g.halo_info(), flush=True)
    for ep in range(5):
        t, v = g.train_epoch(), g.eval(2)
        if rank == 0:
            to, vo = og.train_epoch(), og.eval(2)
            good = abs(t[0] - to[0]) <= 2e-5 * (1 + ep) * abs(to[0]) and abs(v[0] - vo[0]) <= 2e-5 * (1 + ep) * abs(vo[0]) \
                and abs(t[1] - to[1]) < 2e-3 and abs(v[1] - vo[1]) < 4e-3
            ok &= good
            print(name, ep, "native-dist", t, v, "oracle", to, vo, "OK" if good else "MISMATCH", flush=True)
    if rank == 0:
        for l in range(2):
            w = g.weight(l)
            good = np.allclose(w, og.W[l], rtol=2e-4, atol=2e-6)
            ok &= good
            print(name, "W%d" % l, "OK" if good else "MISMATCH", float(np.abs(w - og.W[l]).max()))
    g.close()

# a graph with planted communities (GraphSum goes through the staged kernels on every rank): N ranks vs 1 rank
ds = eng.synth_dataset(40000, 2000000, 64, 7, n_blocks=8, seed=5)
g = eng.GCN(eng.PartDataset(dmod.partition_dataset(ds, rank, world)), comm=comm)
curve = [g.train_epoch() + g.eval(2) for _ in range(3)]
staged = g.timed_epochs(1)["graph_staged"]
g.close()
if rank == 0:
    single = eng.GCN(ds)
    ref = [single.train_epoch() + single.eval(2) for _ in range(3)]
    single.close()
    for ep, (a, b) in enumerate(zip(curve, ref)):
        good = all(abs(x - y) <= 3e-5 * (1 + ep) * max(abs(y), 1e-3) for x, y in zip(a, b))
        ok &= good
        print("synthetic", ep, "staged=%d" % staged, a, b, "OK" if good else "MISMATCH", flush=True)


def against_single_rank(tag, data, env=None):
    """N ranks on `data` (ordinary equal row blocks) against the single-rank engine on the same dataset"""
    global ok
    for k, v in (env or {}).items():
        os.environ[k] = v
    g = eng.GCN(eng.PartDataset(dmod.partition_dataset(data, rank, world)), comm=comm)
    for k in (env or {}):
        del os.environ[k]
    curve = [g.train_epoch() + g.eval(2) for _ in range(3)]
    halo = g.halo_info()
    g.close()
    if rank == 0:
        single = eng.GCN(data)
        ref = [single.train_epoch() + single.eval(2) for _ in range(3)]
        single.close()
        for ep, (a, b) in enumerate(zip(curve, ref)):
            good = all(abs(x - y) <= 3e-5 * (1 + ep) * max(abs(y), 1e-3) for x, y in zip(a, b))
            ok &= good
            print(tag, ep, "halo", halo, a, b, "OK" if good else "MISMATCH", flush=True)
    return halo


# halo exchange forced on: every peer receives only the rows its block references (csrc/comm.cu: p2p_push_rows_kernel)
h = against_single_rank("synthetic-halo", ds, {"GCNB_HALO": "1"})
if world > 1 and comm.gather_mode() == 2 and not h["active"]:
    ok = False
    print("halo exchange was not activated", h, "MISMATCH")
# the same graph with SHUFFLED node ids, laid out by the balanced, community-aligned partitioner (isolated dummy nodes pad
# the ranks' id ranges): an ordinary dataset for the engine; far fewer rows travel than with equal blocks of shuffled ids
shuffled = eng.permute_dataset(ds, np.random.default_rng(99).permutation(ds.num_nodes).astype(np.uint32))
bal, _, info = eng.balanced_partition(shuffled, world)
h_bal = against_single_rank("synthetic-balanced", bal)
h_shuf = against_single_rank("synthetic-shuffled", shuffled)
if rank == 0:
    print("partitioner", {k: info[k] for k in ("communities", "cut_entries", "cut_entries_equal_row_blocks", "rows", "block")},
          "rows needed: balanced", h_bal["rows_needed"], "shuffled equal blocks", h_shuf["rows_needed"], flush=True)
    if world > 1 and not h_bal["rows_needed"] < h_shuf["rows_needed"]:
        ok = False
        print("the balanced partition does not reduce the halo", "MISMATCH")
dist.barrier()
comm.close()
dist.destroy_process_group()
if rank == 0:
    print("DIST_CHECK_NATIVE", "PASS" if ok else "FAIL", "world", world)
    sys.exit(0 if ok else 1)
