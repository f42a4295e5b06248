"""Row-partitioned multi-GPU GCN training: one process per GPU, torch.distributed (NCCL) as plumbing,
every device operation through the C ABI of libgcn_b200 (include/gcnb.h).

Partition: rank r owns the contiguous rows [r*B, min(N,(r+1)*B)) of A_hat, of the features, labels, split and of every
activation (B a multiple of 4 so that a slab always starts on a Philox group boundary).  Weights and Adam state are
replicated.  Communication per pass:
  * one all-gather of the [N x d] GraphSum input per GraphSum (d = 16 with the (A a) W association),
  * one all-reduce of the packed scalars (loss sum, wrong, labelled) per pass,
  * one all-reduce of the flat weight-gradient buffer per training epoch.
Dropout masks come from the stateless Philox keyed by the GLOBAL element index, so the training trajectory does not
depend on the number of ranks (only the fp32 summation order of dW / loss does).

The compute backend is injected (`ops`): `CudaOps` (the product: libgcn_b200 kernels on torch CUDA tensors) or, in the
CPU gloo tests only, a numpy backend the tests build on top of the oracle.  This file never imports the oracle.
"""
import math

import os

import numpy as np


# ---------------------------------------------------------------------------------------------------------------
def block_rows(n, world):
    b = (n + world - 1) // world
    return (b + 3) // 4 * 4


def partition_dataset(ds, rank, world):
    """local slice of a HostDataset-like object (g_indptr, g_indices, f_indptr, f_indices, f_value, label, split,
    optional graph_value).  Column indices stay global."""
    n = len(ds.g_indptr) - 1
    B = block_rows(n, world)
    r0, r1 = min(n, rank * B), min(n, (rank + 1) * B)
    gp = ds.g_indptr.astype(np.int64)
    fp = ds.f_indptr.astype(np.int64)
    deg = np.diff(gp)
    gv = getattr(ds, "graph_value", None)
    if gv is None:
        # 1./sqrtf(deg_src*deg_dst) for the local rows only (src/parser.cpp:164-181); uint32 product like the reference
        rows = np.repeat(np.arange(r0, r1), deg[r0:r1])
        cols = ds.g_indices[gp[r0]:gp[r1]].astype(np.int64)
        prod = (deg[rows].astype(np.uint32) * deg[cols].astype(np.uint32)).astype(np.float32)
        gv_local = (1.0 / np.sqrt(prod, dtype=np.float32).astype(np.float64)).astype(np.float32)
    else:
        gv_local = gv[gp[r0]:gp[r1]]
    out = dict(
        n_global=n, block=B, r0=r0, r1=r1, n_local=r1 - r0,
        g_indptr=(gp[r0:r1 + 1] - gp[r0]).astype(np.uint32), g_indices=np.ascontiguousarray(ds.g_indices[gp[r0]:gp[r1]]),
        graph_value=np.ascontiguousarray(gv_local),
        f_indptr=(fp[r0:r1 + 1] - fp[r0]).astype(np.uint32), f_indices=np.ascontiguousarray(ds.f_indices[fp[r0]:fp[r1]]),
        f_value=np.ascontiguousarray(ds.f_value[fp[r0]:fp[r1]]), f_elem_offset=int(fp[r0]),
        label=np.ascontiguousarray(ds.label[r0:r1]), split=np.ascontiguousarray(ds.split[r0:r1]),
        split_counts=tuple(int((ds.split == s).sum()) for s in (1, 2, 3)),
        f_nnz_global=int(fp[-1]), input_dim=ds.input_dim, output_dim=ds.output_dim)
    return out


# ---------------------------------------------------------------------------------------------------------------
class CudaOps:
    """The product backend: torch CUDA tensors as device memory, libgcn_b200 kernels for every operation."""

    def __init__(self, gcnb, device):
        import torch
        self.t, self.b, self.dev = torch, gcnb, device
        self.launches = 0  # CUDA kernels launched through this backend

    # memory
    def f32(self, *shape):
        return self.t.empty(*shape, dtype=self.t.float32, device=self.dev)

    def u8(self, *shape):
        return self.t.empty(*shape, dtype=self.t.uint8, device=self.dev)

    def zeros_f32(self, *shape):
        return self.t.zeros(*shape, dtype=self.t.float32, device=self.dev)

    def upload(self, a):
        a = np.ascontiguousarray(a)
        if a.dtype == np.uint32:
            a = a.view(np.int32)
        return self.t.from_numpy(a).to(self.dev)

    def to_host(self, x):
        return x.detach().cpu().numpy()

    def scalars(self, k):
        return self.t.zeros(k, dtype=self.t.float64, device=self.dev)

    # kernels
    def plan(self, indptr, indices, n_cols):
        pl = self.b.SpmmPlan(indptr, indices, n_cols)
        pl.kernels = 1 + (pl.info()["n_split_rows"] > 0)
        return pl

    def csc(self, indptr, indices, n_cols):
        c = self.b.Csc(indptr, indices, n_cols)
        if c.plan is not None:
            c.plan.kernels = 1 + (c.plan.info()["n_split_rows"] > 0)
        return c

    def spmm(self, plan, values, B, C, dim, perm=None):
        plan.spmm(values, B, C, dim, perm=perm)
        self.launches += getattr(plan, "kernels", 1)

    def matmul_nn(self, A, B, C, m, n, p):
        self.b.matmul_nn(A, B, C, m, n, p)
        self.launches += 1

    def matmul_nt(self, dC, B, dA, m, n, p):
        self.b.matmul_nt(dC, B, dA, m, n, p)
        self.launches += 1

    def matmul_tn(self, A, dC, dB, m, n, p, ws):
        self.b.matmul_tn(A, dC, dB, m, n, p, ws)
        self.launches += 2

    def tn_workspace(self, m, n, p):
        return self.f32((self.b.lib.gcnb_matmul_tn_workspace(m, n, p) + 3) // 4)

    def rng(self, seed, history, elem_offset):
        return self.b.make_rng(seed, history, elem_offset)

    def glorot(self, w, rows, cols, rng):
        self.b.glorot(w, rows, cols, rng)

    def dropout_oop(self, src, dst, p, rng):
        import ctypes as C
        self.b.check(self.b.lib.gcnb_dropout_fwd_oop_f32(self.b.ptr(src), self.b.ptr(dst), None, None, src.numel(), p,
                                                         C.byref(rng), self.b.stream()))
        self.launches += 1

    def relu_dropout_fwd(self, x, mask, p, training, rng):
        self.b.relu_dropout_fwd(x, mask, p, training, rng=rng)
        self.launches += 1

    def relu_dropout_bwd(self, g, mask, p):
        self.b.relu_dropout_bwd(g, mask, p)
        self.launches += 1

    def set_truth(self, truth, split, label, cur):
        self.b.set_truth(truth, split, label, cur)
        self.launches += 1

    def ce_workspace(self, n):
        return self.b.zeroed_workspace(self.b.lib.gcnb_ce_workspace(n), self.dev)

    def softmax_ce(self, logits, grad, truth, n, C_, num_samples, training, result, ws):
        self.b.softmax_ce(logits, grad, truth, n, C_, num_samples, training, result, ws)
        self.launches += 1

    def sumsq_workspace(self, n):
        return self.b.zeroed_workspace(self.b.lib.gcnb_sumsq_workspace(n), self.dev)

    def sumsq(self, w, out, ws):
        self.b.sumsq(w, out, ws)
        self.launches += 1

    def adam(self, tensors, wd, b1, b2, eps, step_size):
        self.b.adam_step(tensors, wd, b1, b2, eps, step_size)
        self.launches += 1

    def i32(self, n):
        return self.t.empty(n, dtype=self.t.int32, device=self.dev)

    def result_to_scalars(self, result, scal):
        """[loss_sum(float), wrong(bits), labelled(bits)] -> float64 scalars on device, no host sync."""
        scal[0] = result[0].double()
        scal[1:3] = result[1:3].view(self.t.int32).double()

    def sync(self):
        self.t.cuda.synchronize()


class Comm:
    """torch.distributed plumbing (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, dist, rank, world):
        self.d, self.rank, self.world = dist, rank, world

    def all_gather_rows(self, full, local):
        """full: [world*B, d] tensor; local: [B, d] slab (padded to the block size)."""
        if self.world == 1:
            full.copy_(local)
        else:
            self.d.all_gather_into_tensor(full, local)

    def all_reduce(self, x):
        if self.world > 1:
            self.d.all_reduce(x)

    def barrier(self):
        if self.world > 1:
            self.d.barrier()


# ---------------------------------------------------------------------------------------------------------------
class DistGCN:
    """L-layer GCN on one row block.  Mirrors the single-GPU driver (host/src/gcn.cpp): same kernels, same association
    rule, same RNG bookkeeping; GraphSum inputs are all-gathered, weight gradients all-reduced."""

    def __init__(self, part, ops, comm, hidden_dims=(16,), dropouts=(0.5, 0.5), lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8,
                 weight_decay=5e-4, seed=19990304, reorder=True):
        self.p, self.ops, self.comm = part, ops, comm
        o = ops
        self.N, self.B, self.nl = part["n_global"], part["block"], part["n_local"]
        self.dims = [part["input_dim"]] + [int(h) for h in hidden_dims] + [part["output_dim"]]
        self.L = len(self.dims) - 1
        self.dropouts = [float(x) for x in dropouts]
        self.lr, self.b1, self.b2, self.eps, self.wd, self.seed = lr, beta1, beta2, eps, weight_decay, seed
        self.reorder = [False] + [reorder and self.dims[l] < self.dims[l + 1] for l in range(1, self.L)]
        self.step = 0
        self.rng_hist = {}  # n_elements (global) -> times consumed
        # device data
        self.g_indptr, self.g_indices = o.upload(part["g_indptr"]), o.upload(part["g_indices"])
        self.g_value = o.upload(part["graph_value"])
        self.f_indptr, self.f_indices = o.upload(part["f_indptr"]), o.upload(part["f_indices"])
        self.f_value = o.upload(part["f_value"])
        self.label, self.split = o.upload(part["label"]), o.upload(part["split"])
        self.truth = o.i32(max(1, self.nl))
        self.graph_plan = o.plan(self.g_indptr, self.g_indices, self.N)
        if hasattr(self.graph_plan, "stage") and self.nl and 16 in self.dims[1:]:
            # static graph_value: window-staged GraphSum at width 16 (csrc/spmm_stage.cu); no-op without locality
            self.graph_plan.stage(self.g_value, 16, np.ascontiguousarray(part["g_indptr"], np.uint32),
                                  np.ascontiguousarray(part["g_indices"], np.uint32))
        self.feat_csc = o.csc(self.f_indptr, self.f_indices, self.dims[0])
        self.feat_dense = self.feat_csc.is_dense
        self.feat_plan = None if self.feat_dense else o.plan(self.f_indptr, self.f_indices, self.dims[0])
        self.x_drop = o.f32(max(1, part["f_value"].size))
        # replicated weights (Glorot through the global Philox streams: identical on every rank)
        self.W, self.dW, self.m, self.v = [], [], [], []
        sizes = [self.dims[l] * self.dims[l + 1] for l in range(self.L)]
        self.dW_flat = o.zeros_f32(sum(sizes))
        off = 0
        for l in range(self.L):
            w = o.f32(sizes[l])
            o.glorot(w, self.dims[l], self.dims[l + 1], self._rng(0))
            self._consume(sizes[l])
            self.W.append(w)
            self.dW.append(self.dW_flat[off:off + sizes[l]])
            off += sizes[l]
            self.m.append(o.zeros_f32(sizes[l]))
            self.v.append(o.zeros_f32(sizes[l]))
        # per-layer buffers: local slabs padded to the block size (rows beyond n_local stay zero), full gathers
        W_ = comm.world
        self.pre, self.pre_g, self.z, self.z_g, self.mask = [], [], [], [], []
        for l in range(self.L):
            dpre = self.dims[l] if self.reorder[l] else self.dims[l + 1]
            self.pre.append(o.zeros_f32(self.B, dpre))
            self.pre_g.append(o.zeros_f32(self.B, dpre))
            self.z.append(o.zeros_f32(self.B, self.dims[l + 1]))
            self.z_g.append(o.zeros_f32(self.B, self.dims[l + 1]))
            self.mask.append(o.u8(self.B * self.dims[l + 1]) if l + 1 < self.L else None)
        dmax = max(self.dims[1:])
        self.full = o.zeros_f32(W_ * self.B * dmax)  # all-gather target, viewed as [W*B, d] per use
        self.tn_ws = o.tn_workspace(max(1, self.nl), max(self.dims[:-1]), max(self.dims[1:]))
        self.ce_ws = o.ce_workspace(self.nl)
        self.sq_ws = o.sumsq_workspace(sizes[0])
        self.result = o.zeros_f32(4)
        self.l2 = o.zeros_f32(1)
        self.scal = o.scalars(3)
        self.x_train = None

    # -- RNG bookkeeping (global element counts, identical on all ranks) -------------------------------------------
    def _rng(self, elem_offset):
        return self.ops.rng(self.seed, list(self.rng_hist.items()), elem_offset)

    def _consume(self, n_global_elements):
        self.rng_hist[n_global_elements] = self.rng_hist.get(n_global_elements, 0) + 1

    def _gather(self, local_padded, d):
        full = self.full[: self.comm.world * self.B * d].view(self.comm.world * self.B, d)
        self.comm.all_gather_rows(full, local_padded)
        return full

    def _graphsum(self, x_local_padded, out_local_padded, d):
        """out[rows of this rank] = A_hat[rows, :] * all_gather(x)."""
        full = self._gather(x_local_padded, d)
        if self.nl:
            self.ops.spmm(self.graph_plan, self.g_value, full, out_local_padded, d)

    # -- passes ------------------------------------------------------------------------------------------------------
    def forward(self, split, training):
        o, p = self.ops, self.p
        nl, N = self.nl, self.N
        o.set_truth(self.truth, self.split, self.label, split)
        ns = p["split_counts"][split - 1]
        x = self.f_value
        if training:
            p0 = self.dropouts[0]
            if p0 > 0 and nl:
                o.dropout_oop(self.f_value, self.x_drop, p0, self._rng(p["f_elem_offset"]))
                x = self.x_drop
            self._consume(p["f_nnz_global"])
            self.x_train = x
        h0 = self.pre[0]
        if nl:
            if self.feat_dense:
                o.matmul_nn(x, self.W[0], h0, nl, self.dims[0], self.dims[1])
            else:
                o.spmm(self.feat_plan, x, self.W[0], h0, self.dims[1])
        self._graphsum(h0, self.z[0], self.dims[1])
        for l in range(self.L):
            din, dout = self.dims[l], self.dims[l + 1]
            if l > 0:
                a = self.z[l - 1]
                if self.reorder[l]:
                    self._graphsum(a, self.pre[l], din)
                    if nl:
                        o.matmul_nn(self.pre[l], self.W[l], self.z[l], nl, din, dout)
                else:
                    if nl:
                        o.matmul_nn(a, self.W[l], self.pre[l], nl, din, dout)
                    self._graphsum(self.pre[l], self.z[l], dout)
            if l + 1 < self.L:
                if nl:
                    o.relu_dropout_fwd(self.z[l][:nl].view(-1), self.mask[l], self.dropouts[l + 1], training,
                                       self._rng(p["r0"] * dout))
                if training:
                    self._consume(N * dout)
        C_ = self.dims[-1]
        o.softmax_ce(self.z[-1], self.z_g[-1], self.truth, nl, C_, ns, training, self.result, self.ce_ws)
        o.sumsq(self.W[0], self.l2, self.sq_ws)
        o.result_to_scalars(self.result, self.scal)
        self.comm.all_reduce(self.scal)
        self._ns = ns

    def finalize(self):
        s = self.ops.to_host(self.scal)
        l2 = float(self.ops.to_host(self.l2)[0])
        total = self._ns
        loss = np.float32(np.float32(s[0]) / np.float32(total)) + np.float32(np.float32(self.wd) * np.float32(l2) / np.float32(2))
        acc = np.float32(np.float32(total - int(s[1])) / np.float32(total))
        return float(loss), float(acc)

    def backward(self):
        o, nl = self.ops, self.nl
        g = self.z_g[-1]
        for l in range(self.L - 1, 0, -1):
            din, dout = self.dims[l], self.dims[l + 1]
            if self.reorder[l]:
                if nl:
                    o.matmul_tn(self.pre[l], g, self.dW[l], nl, din, dout, self.tn_ws)
                    o.matmul_nt(g, self.W[l], self.pre_g[l], nl, din, dout)
                else:
                    self.dW[l].zero_()
                self._graphsum(self.pre_g[l], self.z_g[l - 1], din)
            else:
                self._graphsum(g, self.pre_g[l], dout)
                if nl:
                    o.matmul_tn(self.z[l - 1], self.pre_g[l], self.dW[l], nl, din, dout, self.tn_ws)
                    o.matmul_nt(self.pre_g[l], self.W[l], self.z_g[l - 1], nl, din, dout)
                else:
                    self.dW[l].zero_()
            if nl:
                o.relu_dropout_bwd(self.z_g[l - 1][:nl].view(-1), self.mask[l - 1], self.dropouts[l])
            g = self.z_g[l - 1]
        self._graphsum(g, self.pre_g[0], self.dims[1])
        if nl:
            if self.feat_dense:
                o.matmul_tn(self.x_train, self.pre_g[0], self.dW[0], nl, self.dims[0], self.dims[1], self.tn_ws)
            else:
                o.spmm(self.feat_csc.plan, self.x_train, self.pre_g[0], self.dW[0], self.dims[1], perm=self.feat_csc.perm)
        else:
            self.dW[0].zero_()
        self.comm.all_reduce(self.dW_flat)  # replicated weights: one all-reduce of all weight gradients per epoch
        self.step += 1
        f32 = np.float32
        step_size = f32(self.lr) * f32(math.sqrt(f32(1) - f32(np.power(f32(self.b2), f32(self.step))))) / \
            (f32(1) - f32(np.power(f32(self.b1), f32(self.step))))
        o.adam([(self.W[l], self.dW[l], self.m[l], self.v[l], l == 0) for l in range(self.L)], self.wd, self.b1, self.b2,
               self.eps, float(step_size))

    def train_epoch(self):
        self.forward(1, True)
        self.backward()
        return self.finalize()

    def eval(self, split):
        self.forward(split, False)
        return self.finalize()


# ---------------------------------------------------------------------------------------------------------------
def make_comm(eng, dist, rank, world, dev):
    """gcnb_comm for this rank: rank 0 draws the NCCL unique id, torch.distributed ships it (plumbing only)."""
    import torch
    ident = torch.zeros(eng.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        ident.copy_(torch.frombuffer(bytearray(eng.Comm.unique_id()), dtype=torch.uint8))
    if world > 1:
        dist.broadcast(ident, 0)
    return eng.Comm(rank, world, bytes(ident.cpu().numpy().tobytes()))


def scaleout_record(rank, world, local_rank, dist, n=4000000, block=4000, intra=230.0, inter=58.0, reflect=2048, sigma=1.0,
                    features=128, classes=47, hidden=16, steps=5, warmup=2, seed=20240229, inter_window=0):
    """Scale-out workload (BASELINE.json configs[4]; north_star: >= 5x 1-GPU throughput on 8 GPUs for a >= 1B-edge synthetic
    graph): a symmetric community graph with 1.01e9 CSR entries generated ROW-LOCALLY (every rank builds only its row block,
    gcnb_synth_sym_rows), dense features, 2-layer GCN hidden 16, through the native (row-partitioned) engine.  Returns the
    record on rank 0 (None elsewhere): ms per train_epoch + eval(2), CUDA events on the engine stream, max over ranks.
    `dist`: an initialised torch.distributed (world > 1) or None."""
    import time
    import torch
    from . import engine as eng
    dev = torch.device("cuda", local_rank)
    B = block_rows(n, world)
    r0, r1 = min(n, rank * B), min(n, (rank + 1) * B)
    rows = r1 - r0
    t0 = time.perf_counter()
    g_indptr, g_indices = eng.synth_sym_rows(n, r0, rows, block, intra, inter, reflect, sigma, seed, inter_window=inter_window)
    t_graph = time.perf_counter() - t0
    deg_local = np.diff(g_indptr.astype(np.int64)).astype(np.uint32)
    if world > 1:
        pad = torch.zeros(B, dtype=torch.int32, device=dev)
        pad[:rows] = torch.from_numpy(deg_local.view(np.int32)).to(dev)
        allpad = torch.empty(world * B, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(allpad, pad)
        deg_global = allpad[:n].cpu().numpy().view(np.uint32)
        nnz_t = torch.tensor([len(g_indices)], dtype=torch.int64, device=dev)
        dist.all_reduce(nnz_t)
        nnz_global = int(nnz_t.item())
    else:
        deg_global, nnz_global = deg_local, len(g_indices)
    gv = eng.synth_graph_values(g_indptr, g_indices, r0, deg_global)
    f_indptr, f_indices, f_value = eng.synth_dense_features_uniform(rows, features, seed, r0 * features)
    label_all, split_all = eng.synth_labels(n, classes, seed=seed)
    t_gen = time.perf_counter() - t0
    part = dict(n_global=n, block=B, r0=r0, r1=r1, n_local=rows, g_indptr=g_indptr, g_indices=g_indices, graph_value=gv,
                f_indptr=f_indptr, f_indices=f_indices, f_value=f_value, f_elem_offset=r0 * features,
                label=np.ascontiguousarray(label_all[r0:r1]), split=np.ascontiguousarray(split_all[r0:r1]),
                f_nnz_global=n * features, input_dim=features, output_dim=classes)
    model = dict(hidden_dims=(hidden,), dropouts=(0.5, 0.5), lr=0.01, weight_decay=5e-4, seed=seed)
    comm = None
    t0 = time.perf_counter()
    if world > 1:
        comm = make_comm(eng, dist, rank, world, dev)
        g = eng.GCN(eng.PartDataset(part), comm=comm, **model)
    else:
        g = eng.GCN(eng.PartDataset(part), **model)
    g.finish_setup()
    torch.cuda.synchronize()
    t_create = time.perf_counter() - t0
    last = None
    for _ in range(warmup):
        last = (g.train_epoch(), g.eval(2))
    clocks = None
    if rank == 0:
        try:
            import bench as _bench
            clocks = _bench.ClockSampler(local_rank).start()
        except Exception:
            clocks = None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    r = g.timed_epochs(steps, with_eval=True, time_graphsum=True)
    torch.cuda.synchronize()
    clk = clocks.stop() if clocks is not None else None
    ms = torch.tensor([r["ms"], r["graphsum_ms"] / max(1, r["graphsum_calls"]), t_create * 1e3, t_gen * 1e3, t_graph * 1e3,
                       r.get("graphsum_exchange_ms", 0.0) / max(1, r["graphsum_calls"])], dtype=torch.float64, device=dev)
    halo = g.halo_info() if world > 1 else None
    halo_sum = torch.tensor([halo["active"], halo["rows_sent"], halo["rows_full_push"], halo["rows_needed"]] if halo else [0, 0, 0, 0],
                            dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(halo_sum)
    line = None
    if rank == 0:
        free_b, total_b = torch.cuda.mem_get_info()
        line = {"workload": "scaleout_sym_community n=%d nnz=%d (community %d, intra %.0f, inter %.0f, sigma %.1f) f=%d c=%d; "
                            "2-layer GCN hidden %d; step = train_epoch + eval(2)" %
                            (n, nnz_global, block, intra, inter, sigma, features, classes, hidden),
                "n_gpus": world, "ms_per_epoch": float(ms[0]) / steps, "unit": "ms/epoch", "scaling": "strong",
                "graphsum_mean_ms": float(ms[1]), "graphsum_calls_per_step": r["graphsum_calls"] / steps,
                "launches_per_step": r["launches"] / steps, "paths": g.path_info(), "setup_ms": float(ms[2]),
                "gen_ms": float(ms[3]), "graph_gen_ms": float(ms[4]), "deg_max": int(deg_global.max()),
                "deg_mean": float(nnz_global / n), "train": last[0] if last else None, "val": last[1] if last else None,
                "gpu_mem_used_gb_rank0": (total_b - free_b) / 2**30, "steps": steps, "warmup": warmup, "clocks": clk,
                "host_cores": os.cpu_count(), "inter_window": inter_window, "graphsum_exchange_mean_ms": float(ms[5]),
                "halo": None if world == 1 else {
                    "ranks_active": int(halo_sum[0]), "rows_sent_per_exchange_all_ranks": int(halo_sum[1]),
                    "rows_of_full_slab_push_all_ranks": int(halo_sum[2]), "rows_needed_all_ranks": int(halo_sum[3]),
                    "fraction_of_full_push": float(halo_sum[1] / max(1.0, float(halo_sum[2])))}}
    g.close()
    if comm is not None:
        comm.close()
    return line


def pin_arrays(part):
    """moves the large numpy arrays of a partition dict into pinned host memory (in place); returns the owning tensors"""
    import torch
    keep = []
    for k, v in list(part.items()):
        if isinstance(v, np.ndarray) and v.nbytes >= (1 << 20):
            a = np.ascontiguousarray(v)
            t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)
            t.numpy()[:] = a.view(np.uint8).reshape(-1)
            part[k] = t.numpy().view(a.dtype).reshape(a.shape)
            keep.append(t)
    return keep


def bench_main(args, rank, world, local_rank, bench):
    """bench.py --gpus N (N > 1): strong scaling of the Reddit-shape epoch.  One rank per GPU; the native engine
    (host/src/gcn.cpp) runs the row block and issues the NCCL collectives itself; torch.distributed only bootstraps
    the communicator and reduces the timings."""
    import json
    import time
    import torch
    import torch.distributed as dist
    from . import engine as eng

    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    ds, w, gen_s = bench.make_dataset(eng, args.scale, pinned=False)
    part = partition_dataset(ds, rank, world)
    nnz_global, n = len(ds.g_indices), ds.num_nodes
    del ds
    pinned = pin_arrays(part)  # the rank's block in pinned host memory, as the single-GPU arm's dataset is  # noqa: F841
    comm = make_comm(eng, dist, rank, world, dev)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    g = eng.GCN(eng.PartDataset(part), hidden_dims=bench.MODEL["hidden"], dropouts=bench.MODEL["dropouts"],
                lr=bench.MODEL["lr"], weight_decay=bench.MODEL["weight_decay"], seed=w["seed"], comm=comm)
    torch.cuda.synchronize()
    t_create = time.perf_counter() - t0
    # e2e leg: K steps through the public calls, every pass reads its loss / accuracy back on the host
    dist.barrier(); torch.cuda.synchronize()
    tw0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        tl = g.train_epoch()
        vl = g.eval(2)
        last = (tl, vl)
    torch.cuda.synchronize(); dist.barrier()
    wall = time.perf_counter() - tw0
    g.finish_setup()
    for _ in range(args.warmup):
        g.train_epoch(); g.eval(2)
    clocks = bench.ClockSampler(local_rank).start()
    dist.barrier(); torch.cuda.synchronize()
    r = g.timed_epochs(args.steps, with_eval=True, time_graphsum=True)  # CUDA events on the engine stream
    torch.cuda.synchronize(); dist.barrier()
    clk = clocks.stop()
    ms = torch.tensor([r["ms"], wall * 1e3, t_create * 1e3, r["graphsum_ms"] / max(1, r["graphsum_calls"]),
                       r["graphsum_exchange_ms"] / max(1, r["graphsum_calls"])], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        step_ms = float(ms[0]) / args.steps
        h2d = sum(v.nbytes for v in part.values() if isinstance(v, np.ndarray))
        d = bench.MODEL["hidden"][0]
        paths = g.path_info()
        gather_mode = {0: "single rank", 1: "NCCL all-gather", 2: "peer-memory push over NVLink (CUDA IPC stores + "
                       "release/acquire flags)"}[int(eng.lib.gcnb_comm_gather_mode(comm.h))]
        gs_path = "bit tiles (tcgen05) + pattern-only ELL remainder" if paths["graph_bittile"] else (
            "window-staged (own-slab windows overlap the exchange) || generic remainder" if paths["graph_staged"] else "generic")
        gs_us = float(ms[3]) * 1e3
        alg = bench.graphsum_alg_bytes(n, nnz_global, d)
        peak, peak_src = bench.peaks()
        achieved = alg / (gs_us * 1e-6) / 1e9  # whole-job bytes per (slowest rank's) GraphSum time
        line = {"metric": bench.METRIC, "value": step_ms, "unit": bench.UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": False, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "reddit_shape_synthetic n=%d nnz=%d f=%d c=%d; 2-layer GCN hidden %d; step = train_epoch + "
                                       "eval(2)" % (n, nnz_global, w["f"], w["c"], d),
                           "parallelism": "row-partitioned x%d (native engine): per GraphSum one exchange of the [N x 16] input "
                                          "slabs (%s), grouped NCCL all-reduce of the weight gradients per epoch and of the "
                                          "loss/count scalars per pass" % (world, gather_mode),
                           "gather_mode": gather_mode, "paths": paths, "halo_exchange_rank0": g.halo_info(),
                           "switches": {k: os.environ[k] for k in sorted(os.environ) if k.startswith("GCNB_")},
                           "l2_policy": "inputs larger than L2", "dataset_gen_s": round(gen_s, 1), "scale": args.scale,
                           "final_train_loss": last[0][0], "final_val_acc": last[1][1]},
                "clocks": clk,
                "e2e": {"value": (float(ms[2]) + float(ms[1])) / args.steps, "unit": bench.UNIT,
                        "h2d_bytes_per_step": int(h2d / args.steps), "d2h_bytes_per_step": 2 * 32,
                        "setup_ms": float(ms[2]),
                        "note": "per rank: upload of its row block + plans, then K steps with per-pass host read of the metrics; "
                                "max over ranks, wall clock / K"},
                "gpu_launches": r["launches"],
                "roofline": {"bound": "hbm", "kernel": "GraphSum d=%d on a row block: slab exchange + %s" % (d, gs_path),
                             "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                             "traffic": None, "peak_source": peak_src + " x %d GPUs" % world,
                             "algorithmic_bytes_per_launch": alg, "mean_launch_us": gs_us,
                             "exchange_us": float(ms[4]) * 1e3, "product_us": gs_us - float(ms[4]) * 1e3,
                             "phases": "exchange_us = start of the call until every peer's slab has landed (max over ranks; with "
                                       "window staging the own-slab windows run inside it), product_us = the rest",
                             "graph_staged": r.get("graph_staged")}}
    g.close()
    comm.close()
    if not getattr(args, "no_scaleout", False):
        rec = scaleout_record(rank, world, local_rank, dist, steps=max(3, min(args.steps, 5)), warmup=2)
        if rank == 0:
            line["scaleout"] = rec
    if rank == 0:
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def bench_main_python_driver(args, rank, world, local_rank, bench):
    """the same measurement through the Python row-partitioned driver (DistGCN); kept for A/B and debugging."""
    import json
    import time
    import torch
    import torch.distributed as dist
    from . import binding as gcnb
    from . import engine as eng

    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    ds, w, gen_s = bench.make_dataset(eng, args.scale, pinned=False)
    part = partition_dataset(ds, rank, world)
    nnz_global, n = len(ds.g_indices), ds.num_nodes
    del ds
    ops, comm = CudaOps(gcnb, dev), Comm(dist, rank, world)
    t0 = time.perf_counter()
    g = DistGCN(part, ops, comm, hidden_dims=bench.MODEL["hidden"], dropouts=bench.MODEL["dropouts"], lr=bench.MODEL["lr"],
                weight_decay=bench.MODEL["weight_decay"], seed=w["seed"])
    torch.cuda.synchronize()
    t_create = time.perf_counter() - t0
    last = None
    for _ in range(args.warmup):
        g.train_epoch(); g.eval(2)
    clocks = bench.ClockSampler(local_rank).start()
    comm.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ops.launches
    tw0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        tl = g.train_epoch()
        vl = g.eval(2)
        last = (tl, vl)
    e1.record()
    comm.barrier(); torch.cuda.synchronize()
    wall = time.perf_counter() - tw0
    clk = clocks.stop()
    ms = torch.tensor([e0.elapsed_time(e1), wall * 1e3, t_create * 1e3], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        step_ms = float(ms[0]) / args.steps
        h2d = sum(v.nbytes for v in part.values() if isinstance(v, np.ndarray))
        line = {"metric": bench.METRIC, "value": step_ms, "unit": bench.UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": False, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "reddit_shape_synthetic n=%d nnz=%d f=%d c=%d; 2-layer GCN hidden %d; step = train_epoch + "
                                       "eval(2)" % (n, nnz_global, w["f"], w["c"], bench.MODEL["hidden"][0]),
                           "parallelism": "row-partitioned x%d: all-gather of the [N x 16] GraphSum input per GraphSum (NCCL), "
                                          "all-reduce of weight gradients per epoch" % world,
                           "l2_policy": "inputs larger than L2", "dataset_gen_s": round(gen_s, 1), "scale": args.scale,
                           "final_train_loss": last[0][0], "final_val_acc": last[1][1]},
                "clocks": clk,
                "e2e": {"value": (float(ms[2]) + float(ms[1])) / args.steps, "unit": bench.UNIT,
                        "h2d_bytes_per_step": int(h2d / args.steps), "d2h_bytes_per_step": 2 * 32,
                        "setup_ms": float(ms[2]),
                        "note": "per rank: upload of its row block + plans, then K steps with per-pass host read of the metrics; "
                                "max over ranks, wall clock / K"},
                "gpu_launches": ops.launches - launches0}
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()
