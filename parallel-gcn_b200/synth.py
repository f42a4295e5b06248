"""Synthetic workloads of BASELINE.json configs 3-5 (host/src/synth.cpp) through their own small host library.

libgcn_synth.so holds ONLY the generator (plain C++, no CUDA, nothing of the product's compute), so that bench.py's
reference arm and the CPU baselines can build the bench graph without mapping libgcn_b200.so.  The same functions are
also exported by libgcn_b200.so for C callers of include/gcnb_engine.h; engine.py re-exports the Python side from here.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgcn_synth.so")
if not os.path.exists(LIB_PATH):
    raise ImportError("libgcn_synth.so is not built (%s): run `python parallel-gcn_b200/build.py`" % LIB_PATH)
lib = C.CDLL(LIB_PATH)
P, I64, I32 = C.c_void_p, C.c_int64, C.c_int32


def _sig(name, res, args):
    fn = getattr(lib, name)
    fn.restype = res
    fn.argtypes = args


_sig("gcnb_synth_graph", I32, [I64, I64, I32, C.c_double, C.c_double, I64, C.c_uint64, P, P, P])
_sig("gcnb_host_free", None, [P])
_sig("gcnb_synth_dense_features", I32, [I64, I32, C.c_uint64, P, P, P])
_sig("gcnb_synth_labels", I32, [I64, I32, C.c_double, C.c_double, C.c_uint64, P, P])
_sig("gcnb_synth_sym_rows", I32, [I64, I64, I64, I64, C.c_double, C.c_double, I32, C.c_double, C.c_uint64, P, P, P])
_sig("gcnb_synth_sym_rows_local", I32, [I64, I64, I64, I64, C.c_double, C.c_double, I32, I64, C.c_double, C.c_uint64, P, P, P])
_sig("gcnb_synth_graph_values", I32, [P, P, I64, I64, P, P])
_sig("gcnb_synth_dense_features_uniform", I32, [I64, I32, C.c_uint64, C.c_uint64, P, P, P])


def check(rc):
    if rc != 0:
        raise RuntimeError("synthetic generator failed with status %d" % rc)


def _p(a):
    return None if a is None else a.ctypes.data_as(P)


class HostDataset:
    """Host CSR of one dataset (uint32 indices, int32 labels) -- the content of the reference's GCNData."""

    FIELDS = ("g_indptr", "g_indices", "f_indptr", "f_indices", "f_value", "label", "split", "graph_value")

    def __init__(self, **kw):
        self.graph_value = None
        self.__dict__.update(kw)

    @property
    def num_nodes(self):
        return len(self.g_indptr) - 1

    def nbytes(self):
        return sum(getattr(self, k).nbytes for k in self.FIELDS if getattr(self, k) is not None)


def synth_graph(n, n_undirected_edges, n_blocks=50, intra=0.8, sigma=1.2, max_deg=21657, seed=19990304):
    ip, ix, nnz = P(), P(), I64(0)
    check(lib.gcnb_synth_graph(n, n_undirected_edges, n_blocks, intra, sigma, max_deg, seed, C.byref(ip), C.byref(ix),
                               C.byref(nnz)))
    indptr = np.ctypeslib.as_array(C.cast(ip, C.POINTER(C.c_uint32)), shape=(n + 1,)).copy()
    indices = np.ctypeslib.as_array(C.cast(ix, C.POINTER(C.c_uint32)), shape=(nnz.value,)).copy()
    lib.gcnb_host_free(ip)
    lib.gcnb_host_free(ix)
    return indptr, indices


def synth_dataset(n, n_undirected_edges, n_features, n_classes, n_blocks=50, intra=0.8, sigma=1.2, max_deg=21657,
                  frac_train=0.66, frac_val=0.10, seed=19990304, pinned=False):
    """Reddit-shape style synthetic dataset (BASELINE.json config 3): planted-community graph, dense N(0,1) features
    stored as an all-columns CSR (how svmlight-Reddit parses), uniform labels, 66/10/24 split."""
    g_indptr, g_indices = synth_graph(n, n_undirected_edges, n_blocks, intra, sigma, max_deg, seed)

    def alloc(shape, dtype):
        if pinned:
            import torch
            t = torch.empty(shape, dtype={np.uint32: torch.int32, np.float32: torch.float32, np.int32: torch.int32}[dtype],
                            pin_memory=True)
            return t.numpy().view(dtype), t
        return np.empty(shape, dtype), None

    keep = []
    f_indptr, t = alloc(n + 1, np.uint32); keep.append(t)
    f_indices, t = alloc(n * n_features, np.uint32); keep.append(t)
    f_value, t = alloc(n * n_features, np.float32); keep.append(t)
    check(lib.gcnb_synth_dense_features(n, n_features, seed, _p(f_indptr), _p(f_indices), _p(f_value)))
    label, split = np.empty(n, np.int32), np.empty(n, np.uint32)
    check(lib.gcnb_synth_labels(n, n_classes, frac_train, frac_val, seed, _p(label), _p(split)))
    if pinned:
        gi, t = alloc(g_indptr.shape, np.uint32); gi[:] = g_indptr; keep.append(t); g_indptr = gi
        gx, t = alloc(g_indices.shape, np.uint32); gx[:] = g_indices; keep.append(t); g_indices = gx
    ds = HostDataset(g_indptr=g_indptr, g_indices=g_indices, f_indptr=f_indptr, f_indices=f_indices, f_value=f_value,
                     label=label, split=split, input_dim=n_features, output_dim=n_classes,
                     split_counts=tuple(int((split == s).sum()) for s in (1, 2, 3)))
    ds._pinned_keepalive = keep
    return ds


def _adopt(ptr_, ctype, n):
    """numpy view of a malloc'ed array returned by the library; freed (gcnb_host_free) when the array is collected"""
    import weakref
    if n == 0:
        lib.gcnb_host_free(ptr_)
        return np.empty(0, np.dtype(ctype))
    a = np.ctypeslib.as_array(C.cast(ptr_, C.POINTER(ctype)), shape=(n,))
    weakref.finalize(a, lib.gcnb_host_free, C.c_void_p(ptr_.value))
    return a


def synth_sym_rows(n, row0, rows, block_size=4000, mean_intra=200.0, mean_inter=50.0, n_reflect=2048, sigma=1.0,
                   seed=19990304, inter_window=0):
    """rows [row0, row0 + rows) of the row-local symmetric community graph (gcnb_synth_sym_rows); global column ids.
    inter_window > 0: the edges that leave a community stay within that many rows of it (gcnb_synth_sym_rows_local)"""
    ip, ix, nnz = P(), P(), I64(0)
    check(lib.gcnb_synth_sym_rows_local(n, row0, rows, block_size, mean_intra, mean_inter, n_reflect, int(inter_window), sigma,
                                        seed, C.byref(ip), C.byref(ix), C.byref(nnz)))
    return _adopt(ip, C.c_uint32, rows + 1), _adopt(ix, C.c_uint32, nnz.value)


def synth_graph_values(indptr, indices, row0, deg_global):
    out = np.empty(len(indices), np.float32)
    deg_global = np.ascontiguousarray(deg_global, np.uint32)
    check(lib.gcnb_synth_graph_values(_p(indptr), _p(indices), len(indptr) - 1, row0, _p(deg_global), _p(out)))
    return out


def synth_dense_features_uniform(rows, n_features, seed, elem_offset=0):
    f_indptr = np.empty(rows + 1, np.uint32)
    f_indices = np.empty(rows * n_features, np.uint32)
    f_value = np.empty(rows * n_features, np.float32)
    check(lib.gcnb_synth_dense_features_uniform(rows, n_features, seed, elem_offset, _p(f_indptr), _p(f_indices), _p(f_value)))
    return f_indptr, f_indices, f_value


def synth_labels(n, n_classes, frac_train=0.66, frac_val=0.10, seed=19990304):
    label, split = np.empty(n, np.int32), np.empty(n, np.uint32)
    check(lib.gcnb_synth_labels(n, n_classes, frac_train, frac_val, seed, _p(label), _p(split)))
    return label, split


