"""torchrun --nproc-per-node N scripts/dist_check.py : N-GPU parity of the row-partitioned driver against the oracle
(cora, citeseer) -- every rank runs CUDA kernels on its GPU, rank 0 compares with the single-process oracle run."""
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
gcnb = importlib.import_module("parallel_gcn_b200.binding")
eng = importlib.import_module("parallel_gcn_b200.engine")
dmod = importlib.import_module("parallel_gcn_b200.dist")
from oracle import oracle as O  # noqa: E402  (checker only)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for name in ("cora", "citeseer"):
    ds = eng.parse_dataset(ROOT, name)
    g = dmod.DistGCN(dmod.partition_dataset(ds, rank, world), dmod.CudaOps(gcnb, torch.device("cuda", local)),
                     dmod.Comm(dist, rank, world))
    og = O.OracleGCN(O.parse_dataset(os.path.join(ROOT, "data", name)), flavour="ref_gpu") if rank == 0 else None
    for ep in range(5):
        t, v = g.train_epoch(), g.eval(2)
        if rank == 0:
            to, vo = og.train_epoch(), og.eval(2)
            good = abs(t[0] - to[0]) <= 2e-5 * (1 + ep) * abs(to[0]) and abs(v[0] - vo[0]) <= 2e-5 * (1 + ep) * abs(vo[0]) \
                and abs(t[1] - to[1]) < 2e-3 and abs(v[1] - vo[1]) < 4e-3
            ok &= good
            print(name, ep, "dist", t, v, "oracle", to, vo, "OK" if good else "MISMATCH", flush=True)
    if rank == 0:
        for l in range(2):
            w = g.W[l].cpu().numpy()
            good = np.allclose(w, og.W[l], rtol=2e-4, atol=2e-6)
            ok &= good
            print(name, "W%d" % l, "OK" if good else "MISMATCH", float(np.abs(w - og.W[l]).max()))
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("DIST_CHECK", "PASS" if ok else "FAIL", "world", world)
    sys.exit(0 if ok else 1)
