"""world_size-2 (and 3) gloo tests of the row-partitioned multi-GPU driver on CPU: partitioning, all-gather / all-reduce
choreography and the global-index Philox offsets, with a numpy compute backend built on the oracle (tests only -- the
product backend is CudaOps).  The distributed trajectory must equal the single-process oracle run."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32, u32, i32, u8 = np.float32, np.uint32, np.int32, np.uint8


class NumpyOps:
    """Same interface as parallel_gcn_b200.dist.CudaOps, computed by gcn_oracle.c on torch CPU tensors."""

    def __init__(self, O):
        import torch
        self.t, self.O = torch, O

    def _np(self, x):
        return x.numpy()

    def f32(self, *shape): return self.t.zeros(*shape, dtype=self.t.float32)
    def u8(self, *shape): return self.t.zeros(*shape, dtype=self.t.uint8)
    def zeros_f32(self, *shape): return self.t.zeros(*shape, dtype=self.t.float32)
    def i32(self, n): return self.t.zeros(n, dtype=self.t.int32)
    def scalars(self, k): return self.t.zeros(k, dtype=self.t.float64)
    def to_host(self, x): return x.numpy()
    def sync(self): pass

    def upload(self, a):
        a = np.ascontiguousarray(a)
        if a.dtype == u32:
            a = a.view(i32)
        return self.t.from_numpy(a.copy())

    def plan(self, indptr, indices, n_cols):
        return ("N", self._np(indptr).view(u32), self._np(indices).view(u32), n_cols)

    def csc(self, indptr, indices, n_cols):
        class _C:
            pass
        c = _C()
        c.is_dense, c.perm = False, None
        c.plan = ("T", self._np(indptr).view(u32), self._np(indices).view(u32), n_cols)
        return c

    def spmm(self, plan, values, B, C, dim, perm=None):
        O, kind, ip, ix, ncols = self.O, *plan
        m = len(ip) - 1
        if kind == "N":
            O.lib.orc_spmm(m, dim, O._p(ip), O._p(ix), O._p(self._np(values)), O._p(self._np(B)), O._p(self._np(C)))
        else:
            O.lib.orc_spmm_bwd(m, ncols, dim, O._p(ip), O._p(ix), O._p(self._np(values)), O._p(self._np(B)), O._p(self._np(C)))

    def matmul_nn(self, A, B, C, m, n, p):
        self.O.lib.orc_matmul(m, n, p, self.O._p(self._np(A)), self.O._p(self._np(B)), self.O._p(self._np(C)))

    def matmul_nt(self, dC, B, dA, m, n, p):
        O = self.O
        dummy_a, dummy_bg = np.zeros((m, n), f32), np.zeros((n, p), f32)
        O.lib.orc_matmul_bwd(m, n, p, O._p(dummy_a), O._p(self._np(B)), O._p(self._np(dC)), O._p(self._np(dA)), O._p(dummy_bg))

    def matmul_tn(self, A, dC, dB, m, n, p, ws):
        O = self.O
        dummy_b, dummy_ag = np.zeros((n, p), f32), np.zeros((m, n), f32)
        O.lib.orc_matmul_bwd(m, n, p, O._p(self._np(A)), O._p(dummy_b), O._p(self._np(dC)), O._p(dummy_ag), O._p(self._np(dB)))

    def tn_workspace(self, m, n, p): return self.f32(1)
    def ce_workspace(self, n): return self.f32(1)
    def sumsq_workspace(self, n): return self.f32(1)

    def rng(self, seed, history, elem_offset):
        return (seed, list(history), elem_offset)

    def _draws(self, history, n_groups):
        d = np.zeros(n_groups, u32)
        for size, cnt in history:
            d[: min(n_groups, (size + 3) // 4)] += cnt
        return d

    def _mask(self, n, p, rng):
        seed, hist, off = rng
        tot = off + n
        m = np.empty(tot, u8)
        self.O.lib.orc_dropout_mask_philox(tot, p, seed, self.O._p(self._draws(hist, (tot + 3) // 4)), 1, self.O._p(m))
        return m[off:]

    def glorot(self, w, rows, cols, rng):
        seed, hist, _ = rng
        a = self._np(w)
        self.O.lib.orc_glorot_philox(a.size, rows, cols, seed, self.O._p(self._draws(hist, (a.size + 3) // 4)), 1, self.O._p(a))

    def dropout_oop(self, src, dst, p, rng):
        m = np.ascontiguousarray(self._mask(src.numel(), p, rng))
        d = self._np(dst)
        d[:] = self._np(src)
        self.O.lib.orc_dropout_apply(d.size, self.O._p(d), self.O._p(m), self.O.lib.orc_dropout_scale(p, 1))

    def relu_dropout_fwd(self, x, mask, p, training, rng):
        O, a = self.O, self._np(x)
        rm = np.zeros(a.size, u8)
        O.lib.orc_relu_fwd(a.size, O._p(a), O._p(rm), 1)
        if training:
            dm = np.ascontiguousarray(self._mask(a.size, p, rng))
            O.lib.orc_dropout_apply(a.size, O._p(a), O._p(dm), O.lib.orc_dropout_scale(p, 1))
            self._np(mask)[: a.size] = rm | (dm << 1)

    def relu_dropout_bwd(self, g, mask, p):
        O, a = self.O, self._np(g)
        mk = self._np(mask)[: a.size]
        dm, rm = np.ascontiguousarray((mk >> 1) & 1), np.ascontiguousarray(mk & 1)
        O.lib.orc_dropout_apply(a.size, O._p(a), O._p(dm), O.lib.orc_dropout_scale(p, 1))
        O.lib.orc_relu_bwd(a.size, O._p(a), O._p(rm))

    def set_truth(self, truth, split, label, cur):
        n = split.numel()
        self.O.lib.orc_set_truth(n, self.O._p(self._np(split).view(u32)), self.O._p(self._np(label)), cur, self.O._p(self._np(truth)))

    def softmax_ce(self, logits, grad, truth, n, C_, num_samples, training, result, ws):
        O = self.O
        lg, gr, tr = self._np(logits), self._np(grad), self._np(truth)
        cnt = np.zeros(1, np.int64)
        loss = O.lib.orc_cross_entropy(n, C_, O._p(lg), O._p(tr), O._p(gr), num_samples, int(training), O._p(cnt)) if n else 0.0
        wrong = O.lib.orc_wrong_count(n, C_, O._p(lg), O._p(tr), None) if n else 0
        r = self._np(result)
        r[0] = loss
        r.view(i32)[1] = wrong
        r.view(i32)[2] = int(cnt[0])

    def result_to_scalars(self, result, scal):
        r = self._np(result)
        s = self._np(scal)
        s[0], s[1], s[2] = float(r[0]), float(r.view(i32)[1]), float(r.view(i32)[2])

    def sumsq(self, w, out, ws):
        self._np(out)[0] = self.O.lib.orc_sumsq(w.numel(), self.O._p(self._np(w)))

    def adam(self, tensors, wd, b1, b2, eps, step_size):
        O = self.O
        for w, g, m, v, decay in tensors:
            gc = np.ascontiguousarray(self._np(g))
            O.lib.orc_adam_step(w.numel(), O._p(self._np(w)), O._p(gc), O._p(self._np(m)), O._p(self._np(v)), int(decay), wd, b1,
                                b2, eps, step_size)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, hidden, dropouts, epochs, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.load_package()
    import importlib
    dmod = importlib.import_module("parallel_gcn_b200.dist")
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    ds = O.parse_dataset(os.path.join(ROOT, "data", name))
    part = dmod.partition_dataset(ds, rank, world)
    g = dmod.DistGCN(part, NumpyOps(O), dmod.Comm(dist, rank, world), hidden_dims=hidden, dropouts=dropouts)
    out = []
    for _ in range(epochs):
        out.append((g.train_epoch(), g.eval(2)))
    if rank == 0:
        q.put((out, [w.numpy().copy() for w in g.W]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,name,hidden,dropouts", [(2, "cora", (16,), (0.5, 0.5)), (3, "citeseer", (16,), (0.5, 0.5)),
                                                        (2, "cora", (8, 24), (0.3, 0.0, 0.2))])
def test_row_partitioned_training_matches_single_process(O, world, name, hidden, dropouts):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    epochs = 3
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, hidden, dropouts, epochs, q)) for r in range(world)]
    for p in procs:
        p.start()
    hist, W = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ds = O.parse_dataset(os.path.join(ROOT, "data", name))
    og = O.OracleGCN(ds, hidden_dims=hidden, dropouts=dropouts, flavour="ref_gpu")
    for ep in range(epochs):
        t, v = og.train_epoch(), og.eval(2)
        (dt, dv) = hist[ep]
        assert abs(dt[0] - t[0]) <= 2e-5 * (1 + ep) * abs(t[0]) and abs(dv[0] - v[0]) <= 2e-5 * (1 + ep) * abs(v[0]), (ep, dt, t, dv, v)
        assert abs(dt[1] - t[1]) < 2e-3 and abs(dv[1] - v[1]) < 4e-3
    for l in range(len(hidden) + 1):
        assert np.allclose(W[l], og.W[l], rtol=2e-4, atol=2e-6), l


def test_partition_covers_everything(O):
    import __graft_entry__ as ge
    ge.load_package()
    import importlib
    dmod = importlib.import_module("parallel_gcn_b200.dist")
    ds = O.parse_dataset(os.path.join(ROOT, "data", "citeseer"))
    for world in (1, 2, 3, 8):
        parts = [dmod.partition_dataset(ds, r, world) for r in range(world)]
        assert sum(p["n_local"] for p in parts) == ds.num_nodes
        assert all(p["r0"] % 4 == 0 for p in parts) and all(p["f_elem_offset"] % 4 == 0 or p["n_local"] == 0 for p in parts[:1])
        gi = np.concatenate([p["g_indices"] for p in parts])
        gv = np.concatenate([p["graph_value"] for p in parts])
        assert (gi == ds.g_indices).all() and (gv.view(u32) == ds.graph_values().view(u32)).all()
        assert (np.concatenate([p["f_value"] for p in parts]).view(u32) == ds.f_value.view(u32)).all()
        assert (np.concatenate([p["label"] for p in parts]) == ds.label).all()
