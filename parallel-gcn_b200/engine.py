"""ctypes binding of the engine-level C ABI (include/gcnb_engine.h): Parser, GCN, synthetic workloads.

Host-side only plumbing: numpy arrays in, numpy arrays / floats out.  The compute is libgcn_b200.so's CUDA path; there
is no fallback -- creating a GCN without a usable sm_100 GPU raises GcnbError.
"""
import ctypes as C

import numpy as np

from .binding import GcnbError, check, lib
from .synth import (HostDataset, synth_dataset, synth_dense_features_uniform, synth_graph,  # noqa: F401
                    synth_graph_values, synth_labels, synth_sym_rows)

P, I64, I32, U32, F32 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_float


class GcnConfig(C.Structure):
    _fields_ = [("num_nodes", I64), ("input_dim", I64), ("output_dim", I64), ("n_layers", I32), ("hidden_dims", P),
                ("dropouts", P), ("epochs", U32), ("early_stopping", U32), ("learning_rate", F32), ("beta1", F32),
                ("beta2", F32), ("eps", F32), ("weight_decay", F32), ("seed", U32), ("quiet", I32), ("reorder", I32)]


class GcnData(C.Structure):
    _fields_ = [("graph_indptr", P), ("graph_indices", P), ("graph_nnz", I64), ("graph_value", P), ("feat_indptr", P),
                ("feat_indices", P), ("feat_value", P), ("feat_nnz", I64), ("label", P), ("split", P)]


class GcnPartition(C.Structure):
    _fields_ = [("comm", P), ("n_global", I64), ("row_offset", I64), ("block", I64), ("feat_elem_offset", I64),
                ("feat_nnz_global", I64)]


def _sig(name, res, args):
    fn = getattr(lib, name)
    fn.restype = res
    fn.argtypes = args


_sig("gcnb_dataset_parse", I32, [C.c_char_p, C.c_char_p, I32, P])
_sig("gcnb_dataset_dims", I32, [P, P])
_sig("gcnb_dataset_copy", I32, [P, I32, P])
_sig("gcnb_dataset_free", I32, [P])
_sig("gcnb_dataset_save", I32, [P, C.c_char_p])
_sig("gcnb_dataset_load", I32, [C.c_char_p, P])
_sig("gcnb_gcn_create", I32, [P, P, P])
_sig("gcnb_gcn_create_from_dataset", I32, [P, P, P])
_sig("gcnb_gcn_create_partitioned", I32, [P, P, P, P])
_sig("gcnb_comm_unique_id", I32, [P])
_sig("gcnb_comm_create", I32, [I32, I32, P, P])
_sig("gcnb_comm_destroy", I32, [P])
_sig("gcnb_comm_gather_mode", I32, [P])
_sig("gcnb_gcn_destroy", I32, [P])
_sig("gcnb_gcn_train_epoch", I32, [P, P])
_sig("gcnb_gcn_eval", I32, [P, I32, P])
_sig("gcnb_gcn_run", I32, [P, P])
_sig("gcnb_gcn_weight_size", I64, [P, I32])
_sig("gcnb_gcn_get_weight", I32, [P, I32, P])
_sig("gcnb_gcn_set_weight", I32, [P, I32, P])
_sig("gcnb_gcn_get_weight_grad", I32, [P, I32, P])
_sig("gcnb_gcn_get_logits", I32, [P, P])
_sig("gcnb_gcn_set_mask", I32, [P, I32, P])
_sig("gcnb_gcn_launches_per_epoch", I64, [P])
_sig("gcnb_gcn_graph_staged", I32, [P])
_sig("gcnb_gcn_graph_bittile", I32, [P])
_sig("gcnb_gcn_path_info", I32, [P, P])
_sig("gcnb_gcn_set_cuda_graph", I32, [P, I32])
_sig("gcnb_gcn_finish_setup", I32, [P])
_sig("gcnb_gcn_uses_cuda_graph", I32, [P])
_sig("gcnb_gcn_launches_total", I64, [P])
_sig("gcnb_gcn_timed_epochs", I32, [P, I32, I32, I32, P])
_sig("gcnb_gcn_graphsum_exchange_ms", C.c_double, [P])
_sig("gcnb_host_free", None, [P])
_sig("gcnb_sweep_run", I32, [P, P, I64, I32, P, P])
_sig("gcnb_reorder_communities", I32, [I64, P, P, I32, C.c_uint64, P, P])
_sig("gcnb_partition_communities", I32, [I64, P, P, I32, I32, C.c_uint64, C.c_double, P, P, P, P])
_sig("gcnb_gcn_halo_info", I32, [P, P])
_sig("gcnb_permute_csr", I32, [I64, P, P, P, P, P])
_sig("gcnb_permute_rows", I32, [I64, I64, P, P, P])
_sig("gcnb_unpermute_rows", I32, [I64, I64, P, P, P])


def _p(a):
    return None if a is None else a.ctypes.data_as(P)


class SweepTrial(C.Structure):
    _fields_ = [("n_layers", I32), ("hidden_dims", U32 * 7), ("dropouts", F32 * 8), ("epochs", U32), ("early_stopping", U32),
                ("learning_rate", F32), ("weight_decay", F32), ("seed", U32)]


class SweepResult(C.Structure):
    _fields_ = [("last_val_accuracy", F32), ("last_val_loss", F32), ("last_train_loss", F32), ("avg_epoch_ms", F32),
                ("total_s", F32), ("epochs_run", U32)]


def sweep_run(source, trials, workers=0):
    """The reference's tuning sweep (test/tuning_accuracy.cpp) as a throughput workload: gcnb_sweep_run.

    source: (root, name) of a shipped dataset or the path of a binary container; trials: list of dicts with keys
    hidden_dims, dropouts, epochs, early_stopping, learning_rate, weight_decay, seed.  Returns (list of result dicts in
    trial order, wall seconds).  Results do not depend on `workers`."""
    h = P()
    if isinstance(source, (tuple, list)):
        rc = lib.gcnb_dataset_parse(str(source[0]).encode(), source[1].encode(), 0, C.byref(h))
    else:
        rc = lib.gcnb_dataset_load(str(source).encode(), C.byref(h))
    if rc != 0:
        raise GcnbError("sweep_run: cannot read the dataset %r (code %d)" % (source, rc))
    try:
        arr = (SweepTrial * len(trials))()
        for t, src in zip(arr, trials):
            hd, dr = list(src.get("hidden_dims", [16])), list(src.get("dropouts", [0.5, 0.5]))
            t.n_layers = len(hd) + 1
            assert len(dr) == t.n_layers and t.n_layers <= 8
            for i, v in enumerate(hd):
                t.hidden_dims[i] = int(v)
            for i, v in enumerate(dr):
                t.dropouts[i] = float(v)
            t.epochs, t.early_stopping = int(src.get("epochs", 100)), int(src.get("early_stopping", 0))
            t.learning_rate, t.weight_decay = float(src.get("learning_rate", 0.01)), float(src.get("weight_decay", 5e-4))
            t.seed = int(src.get("seed", 19990304))
        res = (SweepResult * len(trials))()
        wall = C.c_double(0)
        check(lib.gcnb_sweep_run(h, arr, len(trials), int(workers), res, C.byref(wall)))
        out = [{f: getattr(r, f) for f, _ in SweepResult._fields_} for r in res]
        return out, wall.value
    finally:
        lib.gcnb_dataset_free(h)


def parse_dataset(root, name, no_feature=False, save_to=None):
    """Parser(params, data, name).parse() run from `root` (expects root/data/<name>.graph|.split|.svmlight);
    save_to: also store the parsed dataset in the binary container (gcnb_dataset_save)"""
    h = P()
    rc = lib.gcnb_dataset_parse(str(root).encode(), name.encode(), int(no_feature), C.byref(h))
    if rc != 0:
        return None
    if save_to is not None:
        check(lib.gcnb_dataset_save(h, str(save_to).encode()))
    return _dataset_from_handle(h)


def load_dataset(path):
    """a dataset stored by gcnb_dataset_save: bit-identical to parsing the text files again"""
    h = P()
    if lib.gcnb_dataset_load(str(path).encode(), C.byref(h)) != 0:
        return None
    return _dataset_from_handle(h)


def _dataset_from_handle(h):
    dims = (I64 * 10)()
    check(lib.gcnb_dataset_dims(h, dims))
    n, gnnz, frows, fnnz, in_dim, out_dim, nsplit, tr, va, te = (int(x) for x in dims)
    arrs = [np.empty(n + 1, np.uint32), np.empty(gnnz, np.uint32), np.empty(frows + 1, np.uint32),
            np.empty(fnnz, np.uint32), np.empty(fnnz, np.float32), np.empty(frows, np.int32),
            np.empty(nsplit, np.uint32), np.empty(gnnz, np.float32)]
    for k, a in enumerate(arrs):
        check(lib.gcnb_dataset_copy(h, k, _p(a)))
    lib.gcnb_dataset_free(h)
    return HostDataset(**dict(zip(HostDataset.FIELDS, arrs)), input_dim=in_dim, output_dim=out_dim,
                       split_counts=(tr, va, te))


def reorder_communities(indptr, indices, max_sweeps=0, seed=1):
    """new_of_old permutation that makes label-propagation communities contiguous; returns (new_of_old, n_communities)"""
    n = len(indptr) - 1
    new_of_old, nc = np.empty(n, np.uint32), I64(0)
    check(lib.gcnb_reorder_communities(n, _p(indptr), _p(indices), int(max_sweeps), int(seed), _p(new_of_old), C.byref(nc)))
    return new_of_old, int(nc.value)


def partition_communities(indptr, indices, world, max_sweeps=0, seed=1, tolerance=0.0):
    """balanced, community-aligned row partition (gcnb_partition_communities): returns (new_of_old into the padded id space
    [0, world * block), block, rows per rank, stats dict)"""
    n = len(indptr) - 1
    new_of_old, block, rows, stats = np.empty(n, np.uint32), I64(0), (I64 * world)(), (I64 * 4)()
    check(lib.gcnb_partition_communities(n, _p(indptr), _p(indices), int(world), int(max_sweeps), int(seed), float(tolerance),
                                         _p(new_of_old), C.byref(block), rows, stats))
    return new_of_old, int(block.value), [int(r) for r in rows], dict(
        communities=int(stats[0]), cut_entries=int(stats[1]), cut_entries_equal_row_blocks=int(stats[2]),
        max_rank_entries=int(stats[3]), total_entries=int(indptr[-1]))


def pad_dataset(ds, n_pad):
    """the dataset with isolated dummy nodes appended up to n_pad nodes: a self entry in the graph, no label, no split, a
    zero feature row (fixed-width datasets keep their width so that the all-columns detection still holds)"""
    n = ds.num_nodes
    extra = n_pad - n
    assert extra >= 0
    if extra == 0:
        return ds
    g_indptr = np.concatenate([ds.g_indptr, ds.g_indptr[-1] + np.arange(1, extra + 1, dtype=np.uint32)]).astype(np.uint32)
    g_indices = np.concatenate([ds.g_indices, np.arange(n, n_pad, dtype=np.uint32)])
    flen = np.diff(ds.f_indptr.astype(np.int64))
    if len(flen) and (flen == flen[0]).all() and flen[0] > 0:
        w = int(flen[0])
        f_indptr = (np.arange(n_pad + 1, dtype=np.int64) * w).astype(np.uint32)
        f_indices = np.concatenate([ds.f_indices, np.tile(ds.f_indices[:w], extra)])
        f_value = np.concatenate([ds.f_value, np.zeros(extra * w, np.float32)])
    else:
        f_indptr = np.concatenate([ds.f_indptr, np.full(extra, ds.f_indptr[-1], np.uint32)])
        f_indices, f_value = ds.f_indices, ds.f_value
    out = HostDataset(g_indptr=g_indptr, g_indices=g_indices, f_indptr=f_indptr, f_indices=f_indices, f_value=f_value,
                      label=np.concatenate([ds.label, np.full(extra, -1, np.int32)]),
                      split=np.concatenate([ds.split, np.zeros(extra, np.uint32)]), input_dim=ds.input_dim,
                      output_dim=ds.output_dim)
    if getattr(ds, "split_counts", None) is not None:
        out.split_counts = ds.split_counts
    return out


def balanced_partition(ds, world, max_sweeps=0, seed=1, tolerance=0.0):
    """Dataset laid out for `world` ranks by gcnb_partition_communities: rank r's nodes occupy the ids
    [r * block, r * block + rows[r]) of the returned dataset (world * block nodes; the ids in between are isolated dummy nodes
    without label or split), so dist.partition_dataset(out, r, world) hands every rank the same number of CSR entries and
    whole communities.  Returns (dataset, new_of_old, info); per-node outputs go back with permute_rows(..., inverse=True)
    on the first num_nodes entries of new_of_old."""
    new_of_old, block, rows, stats = partition_communities(ds.g_indptr, ds.g_indices, world, max_sweeps, seed, tolerance)
    n_pad = world * block
    used = np.zeros(n_pad, bool)
    used[new_of_old] = True
    full = np.concatenate([new_of_old, np.flatnonzero(~used).astype(np.uint32)])
    out = permute_dataset(pad_dataset(ds, n_pad), full)
    return out, new_of_old, dict(stats, block=block, rows=rows, world=world)


def permute_rows(a, new_of_old, inverse=False):
    """rows of `a` moved to their new positions (inverse: back to the original numbering)"""
    a = np.ascontiguousarray(a)
    out = np.empty_like(a)
    row_bytes = a.nbytes // max(1, a.shape[0])
    fn = lib.gcnb_unpermute_rows if inverse else lib.gcnb_permute_rows
    check(fn(a.shape[0], row_bytes, _p(new_of_old), _p(a), _p(out)))
    return out


def permute_dataset(ds, new_of_old):
    """the dataset renumbered by new_of_old (all-columns / fixed-width feature rows or general CSR features)"""
    n = ds.num_nodes
    g_indptr, g_indices = np.empty(n + 1, np.uint32), np.empty(len(ds.g_indices), np.uint32)
    check(lib.gcnb_permute_csr(n, _p(ds.g_indptr), _p(ds.g_indices), _p(new_of_old), _p(g_indptr), _p(g_indices)))
    flen = np.diff(ds.f_indptr.astype(np.int64))
    if len(flen) and (flen == flen[0]).all() and flen[0] > 0:   # fixed-width rows: one row permutation
        w = int(flen[0])
        f_indices = permute_rows(ds.f_indices.reshape(n, w), new_of_old).reshape(-1)
        f_value = permute_rows(ds.f_value.reshape(n, w), new_of_old).reshape(-1)
        f_indptr = ds.f_indptr.copy()
    else:
        old_of_new = np.empty(n, np.int64)
        old_of_new[new_of_old] = np.arange(n)
        f_indptr = np.zeros(n + 1, np.uint32)
        f_indptr[1:] = np.cumsum(flen[old_of_new])
        src = np.concatenate([np.arange(ds.f_indptr[i], ds.f_indptr[i + 1]) for i in old_of_new]) if n else np.empty(0, np.int64)
        f_indices, f_value = ds.f_indices[src], ds.f_value[src]
    out = HostDataset(g_indptr=g_indptr, g_indices=g_indices, f_indptr=f_indptr, f_indices=np.ascontiguousarray(f_indices),
                      f_value=np.ascontiguousarray(f_value), label=permute_rows(ds.label, new_of_old),
                      split=permute_rows(ds.split, new_of_old), input_dim=ds.input_dim, output_dim=ds.output_dim)
    if getattr(ds, "split_counts", None) is not None:
        out.split_counts = ds.split_counts
    return out


COMM_ID_BYTES = 128


class Comm:
    """gcnb_comm: the NCCL communicator of the row-partitioned engine.  `id_bytes` (rank 0: Comm.unique_id()) has to
    reach every rank through a side channel (bench.py / scripts use a torch.distributed broadcast)."""

    @staticmethod
    def unique_id():
        buf = (C.c_ubyte * COMM_ID_BYTES)()
        check(lib.gcnb_comm_unique_id(buf))
        return bytes(buf)

    def __init__(self, rank, world, id_bytes=None):
        self.rank, self.world = int(rank), int(world)
        h = P()
        buf = None if id_bytes is None else (C.c_ubyte * COMM_ID_BYTES).from_buffer_copy(id_bytes)
        check(lib.gcnb_comm_create(self.rank, self.world, buf, C.byref(h)))
        self.h = h

    def gather_mode(self):
        """0 single rank, 1 NCCL all-gather, 2 peer-memory push (valid once a model has set the exchange up)"""
        return int(lib.gcnb_comm_gather_mode(self.h))

    def close(self):
        if getattr(self, "h", None) and lib is not None:
            lib.gcnb_comm_destroy(self.h)
            self.h = None


class PartDataset(HostDataset):
    """row block of a dataset as produced by dist.partition_dataset, in HostDataset clothing (column ids global)."""

    def __init__(self, part):
        super().__init__(g_indptr=part["g_indptr"], g_indices=part["g_indices"], f_indptr=part["f_indptr"],
                         f_indices=part["f_indices"], f_value=part["f_value"], label=part["label"], split=part["split"],
                         graph_value=part["graph_value"], input_dim=part["input_dim"], output_dim=part["output_dim"])
        self.part = part


class GCN:
    """The reference's GCN driver (GCN(params, adam_params, data); run(); private train_epoch/eval exposed).
    With `comm` (and ds = PartDataset) the model runs on one row block of a row-partitioned multi-GPU job."""

    def __init__(self, ds, hidden_dims=(16,), dropouts=(0.5, 0.5), epochs=100, early_stopping=0, lr=0.01, beta1=0.9,
                 beta2=0.999, eps=1e-8, weight_decay=5e-4, seed=19990304, quiet=True, reorder=True, comm=None):
        self.ds = ds
        self.n_layers = len(hidden_dims) + 1
        assert len(dropouts) == self.n_layers
        self._hd = np.asarray(hidden_dims, np.uint32)
        self._dp = np.asarray(dropouts, np.float32)
        cfg = GcnConfig(ds.num_nodes, ds.input_dim, ds.output_dim, self.n_layers, _p(self._hd) if len(self._hd) else None,
                        _p(self._dp), epochs, early_stopping, lr, beta1, beta2, eps, weight_decay, seed, int(quiet),
                        int(reorder))
        gv = getattr(ds, "graph_value", None)
        data = GcnData(_p(ds.g_indptr), _p(ds.g_indices), len(ds.g_indices), _p(gv), _p(ds.f_indptr), _p(ds.f_indices),
                       _p(ds.f_value), len(ds.f_indices), _p(ds.label), _p(ds.split))
        h = P()
        if comm is None:
            check(lib.gcnb_gcn_create(C.byref(cfg), C.byref(data), C.byref(h)))
        else:
            pt = ds.part
            self.comm = comm  # keep alive: the model borrows it
            part = GcnPartition(comm.h, pt["n_global"], pt["r0"], pt["block"], pt["f_elem_offset"], pt["f_nnz_global"])
            check(lib.gcnb_gcn_create_partitioned(C.byref(cfg), C.byref(data), C.byref(part), C.byref(h)))
        self.h = h
        self.dims = [ds.input_dim] + [int(x) for x in hidden_dims] + [ds.output_dim]

    def _pair(self, fn, *a):
        out = (F32 * 2)()
        check(fn(self.h, *a, out))
        return float(out[0]), float(out[1])

    def train_epoch(self):
        return self._pair(lib.gcnb_gcn_train_epoch)

    def eval(self, split):
        return self._pair(lib.gcnb_gcn_eval, int(split))

    def halo_info(self):
        """row-partitioned models: dict(active, rows_sent, rows_full_push, rows_needed) of the halo exchange"""
        out = (I64 * 4)()
        check(lib.gcnb_gcn_halo_info(self.h, out))
        return dict(active=int(out[0]), rows_sent=int(out[1]), rows_full_push=int(out[2]), rows_needed=int(out[3]))

    def run(self):
        out = (F32 * 4)()
        check(lib.gcnb_gcn_run(self.h, out))
        return dict(avg_epoch_ms=float(out[0]), total_s=float(out[1]), last_val_acc=float(out[2]), epochs=int(out[3]))

    def weight(self, l):
        w = np.empty(lib.gcnb_gcn_weight_size(self.h, l), np.float32)
        check(lib.gcnb_gcn_get_weight(self.h, l, _p(w)))
        return w

    def weight_grad(self, l):
        w = np.empty(lib.gcnb_gcn_weight_size(self.h, l), np.float32)
        check(lib.gcnb_gcn_get_weight_grad(self.h, l, _p(w)))
        return w

    def set_weight(self, l, w):
        w = np.ascontiguousarray(w, np.float32).ravel()
        assert w.size == lib.gcnb_gcn_weight_size(self.h, l)
        check(lib.gcnb_gcn_set_weight(self.h, l, _p(w)))

    def logits(self):
        out = np.empty((self.ds.num_nodes, self.ds.output_dim), np.float32)
        check(lib.gcnb_gcn_get_logits(self.h, _p(out)))
        return out

    def set_mask(self, site, mask):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        check(lib.gcnb_gcn_set_mask(self.h, site, _p(m)))

    def timed_epochs(self, n_epochs, with_eval=True, time_graphsum=False):
        out = (F32 * 4)()
        check(lib.gcnb_gcn_timed_epochs(self.h, int(n_epochs), int(with_eval), int(time_graphsum), out))
        return dict(ms=float(out[0]), graphsum_ms=float(out[1]), graphsum_calls=int(out[2]), launches=int(out[3]),
                    graph_staged=int(lib.gcnb_gcn_graph_staged(self.h)),
                    graphsum_exchange_ms=float(lib.gcnb_gcn_graphsum_exchange_ms(self.h)))

    def launches_per_epoch(self):
        return int(lib.gcnb_gcn_launches_per_epoch(self.h))

    def graph_staged(self):
        return bool(lib.gcnb_gcn_graph_staged(self.h))

    def graph_bittile(self):
        return bool(lib.gcnb_gcn_graph_bittile(self.h))

    def path_info(self):
        """which fast paths are active (gcnb_gcn_path_info)"""
        out = (C.c_int * 8)()
        check(lib.gcnb_gcn_path_info(self.h, out))
        keys = ("graph_staged", "graph_bittile", "dense_fast", "propagated_features", "cuda_graph", "setup_pending",
                "dense_tc")
        d = dict(zip(keys, [bool(x) for x in out]))
        d["partitioned"], d["graph_renumbered"] = out[7] == 1, out[7] == 2
        return d

    def finish_setup(self):
        """attach the background-staged GraphSum representation now (GCNB_ASYNC_STAGE=1); no-op otherwise"""
        check(lib.gcnb_gcn_finish_setup(self.h))

    def set_cuda_graph(self, on):
        check(lib.gcnb_gcn_set_cuda_graph(self.h, int(on)))

    def uses_cuda_graph(self):
        return bool(lib.gcnb_gcn_uses_cuda_graph(self.h))

    def close(self):
        if getattr(self, "h", None) and lib is not None:
            lib.gcnb_gcn_destroy(self.h)
            self.h = None

    __del__ = close
