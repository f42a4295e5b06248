"""ctypes binding of libgcn_b200.so (the C ABI declared in include/gcnb.h).

PyTorch is used only as plumbing: device memory (tensor.data_ptr()) and the current CUDA stream.  There is no
CPU / eager fallback: importing this module without the built library, or calling any kernel without a usable
sm_100 device, raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GCNB_LIB") or os.path.join(HERE, "libgcn_b200.so")  # GCNB_LIB: tuning builds only

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libgcn_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `python parallel-gcn_b200/build.py`; there is no fallback path." % LIB_PATH)

lib = C.CDLL(LIB_PATH)
P, I64, I32, U32, F32 = C.c_void_p, C.c_int64, C.c_int, C.c_uint32, C.c_float

MAX_RNG_HIST = 48  # GCNB_MAX_RNG_HIST / GCNB_MAX_TENSORS of include/gcnb.h
MAX_TENSORS = 48


class RngT(C.Structure):
    _fields_ = [("seed", U32), ("group_offset", U32), ("elem_lead", U32), ("n_hist", I32), ("hist_groups", U32 * MAX_RNG_HIST), ("hist_count", U32 * MAX_RNG_HIST)]


class AdamTensorsT(C.Structure):
    _fields_ = [("n_tensors", I32), ("w", P * MAX_TENSORS), ("g", P * MAX_TENSORS), ("m", P * MAX_TENSORS),
                ("v", P * MAX_TENSORS), ("size", I64 * MAX_TENSORS), ("decay", I32 * MAX_TENSORS)]


def _sig(name, res, args):
    fn = getattr(lib, name)
    fn.restype = res
    fn.argtypes = args
    return fn


_sig("gcnb_error_string", C.c_char_p, [I32])
_sig("gcnb_version", I32, [])
_sig("gcnb_device_check", I32, [P])
_sig("gcnb_spmm_plan_create", I32, [P, P, I64, I64, I32, P, P])
_sig("gcnb_spmm_plan_destroy", I32, [P])
_sig("gcnb_spmm_plan_info", I32, [P, P])
_sig("gcnb_spmm_f32", I32, [P, P, P, P, P, I32, P])
_sig("gcnb_spmm_ld_f32", I32, [P, P, P, P, I64, P, I64, I32, P])
_sig("gcnb_spmm_plan_stage", I32, [P, P, P, P, I32, P])
_sig("gcnb_spmm_plan_stage_ex", I32, [P, P, P, P, I32, I32, I32, I32, I64, P])
_sig("gcnb_spmm_plan_stage_info", I32, [P, P])
_sig("gcnb_spmm_plan_stage_async_begin", I32, [P, P, I32, P])
_sig("gcnb_spmm_plan_stage_async_done", I32, [P])
_sig("gcnb_spmm_plan_stage_async_finish", I32, [P, P])
_sig("gcnb_spmm_plan_stage_slabs", I32, [P, I32])
_sig("gcnb_stage_host_build", I32, [P, P, I64, I64, I32, I32, I32, I32, I64, I32, I32, P])
_sig("gcnb_stage_host_build_own", I32, [P, P, I64, I64, I32, I32, I32, I32, I64, I32, I32, I64, I64, P])
_sig("gcnb_spmm_plan_set_own_cols", I32, [P, I64, I64])
_sig("gcnb_spmm_stage_own_f32", I32, [P, P, P, I32, P, P])
_sig("gcnb_stage_host_sizes", I32, [P, P])
_sig("gcnb_stage_host_copy", I32, [P, I32, P, I64])
_sig("gcnb_stage_host_destroy", I32, [P])
_sig("gcnb_bittile_plan_create", I32, [P, P, P, I64, I64, P, P, I32, I32, I32, P, P])
_sig("gcnb_bittile_plan_destroy", I32, [P])
_sig("gcnb_bittile_plan_create_device", I32, [P, P, P, I64, I64, P, P, I32, I32, I32, P, P])
_sig("gcnb_bittile_plan_sizes", I32, [P, P])
_sig("gcnb_bittile_plan_copy", I32, [P, I32, P, I64])
_sig("gcnb_bittile_plan_info", I32, [P, P])
_sig("gcnb_bittile_spmm16_f32", I32, [P, P, P, P])
_sig("gcnb_bittile_spmm_ld_f32", I32, [P, P, I64, P, I64, I32, P])
_sig("gcnb_bittile_debug_pack", I32, [P, P, P, I64, P])
_sig("gcnb_bittile_debug_parts", I32, [P, I32])
_sig("gcnb_spmm_plan_attach_bittile", I32, [P, P, P])
_sig("gcnb_bittile_host_build", I32, [P, P, P, I64, I64, P, P, I32, I32, I32, I32, I32, P])
_sig("gcnb_bittile_host_sizes", I32, [P, P])
_sig("gcnb_bittile_host_copy", I32, [P, I32, P, I64])
_sig("gcnb_bittile_host_destroy", I32, [P])
_sig("gcnb_ell_host_build", I32, [P, P, I64, I64, I32, P])
_sig("gcnb_ell_host_sizes", I32, [P, P])
_sig("gcnb_ell_host_copy", I32, [P, I32, P, I64])
_sig("gcnb_ell_host_destroy", I32, [P])
_sig("gcnb_ell_plan_create", I32, [P, P, I64, I64, P, P])
_sig("gcnb_ell_plan_destroy", I32, [P])
_sig("gcnb_ell_gather16_f32", I32, [P, P, P, P, P])
_sig("gcnb_csc_create", I32, [P, P, I64, I64, P, P])
_sig("gcnb_csc_destroy", I32, [P])
_sig("gcnb_csc_arrays", I32, [P, P, P, P, P])
_sig("gcnb_matmul_nn_f32", I32, [P, P, P, I64, I32, I32, P])
_sig("gcnb_matmul_nt_f32", I32, [P, P, P, I64, I32, I32, P])
_sig("gcnb_matmul_tn_workspace", I64, [I64, I32, I32])
_sig("gcnb_matmul_tn_f32", I32, [P, P, P, I64, I32, I32, P, I64, P])
_sig("gcnb_dense_feat_supported", I32, [I32, I32])
_sig("gcnb_dropout_maskbits_words", I64, [I64, I32])
_sig("gcnb_dropout_maskbits", I32, [P, I64, I32, F32, P, P])
_sig("gcnb_dense_feat_fwd_f32", I32, [P, P, F32, P, P, I64, I32, I32, P])
_sig("gcnb_dense_feat_tn_workspace", I64, [I64, I32, I32])
_sig("gcnb_dense_feat_tn_f32", I32, [P, P, F32, P, P, I64, I32, I32, P, I64, P])
_sig("gcnb_dense_tc_supported", I32, [I32, I32])
_sig("gcnb_dense_tc_x_bytes", I64, [I64, I32])
_sig("gcnb_dense_tc_w_bytes", I64, [I32, I32])
_sig("gcnb_dense_tc_pack_x", I32, [P, P, I64, I32, P])
_sig("gcnb_dense_tc_fwd_f32", I32, [P, P, P, I64, I32, I32, P, I64, P])
_sig("gcnb_dense_tc_xt_bytes", I64, [I64, I32])
_sig("gcnb_dense_tc_pack_xt", I32, [P, P, I64, I32, P])
_sig("gcnb_dense_tc_tn_workspace", I64, [I64, I32, I32])
_sig("gcnb_dense_tc_tn_f32", I32, [P, P, P, I64, I32, I32, P, I64, P])
_sig("gcnb_glorot_f32", I32, [P, I64, U32, U32, P, P])
_sig("gcnb_dropout_fwd_f32", I32, [P, P, P, I64, F32, P, P])
_sig("gcnb_dropout_fwd_oop_f32", I32, [P, P, P, P, I64, F32, P, P])
_sig("gcnb_dropout_bwd_f32", I32, [P, P, I64, F32, P])
_sig("gcnb_relu_fwd_f32", I32, [P, P, I64, I32, P])
_sig("gcnb_relu_bwd_f32", I32, [P, P, I64, P])
_sig("gcnb_relu_dropout_fwd_f32", I32, [P, P, P, I64, F32, I32, P, P])
_sig("gcnb_relu_dropout_bwd_f32", I32, [P, P, I64, F32, P])
_sig("gcnb_set_truth", I32, [P, P, P, I64, U32, P])
_sig("gcnb_graph_values_f32", I32, [P, P, I64, P, P])
_sig("gcnb_ce_workspace", I64, [I64])
_sig("gcnb_softmax_ce_f32", I32, [P, P, P, I64, I32, U32, I32, P, P, P])
_sig("gcnb_head_supported", I32, [I32, I32])
_sig("gcnb_head_workspace", I64, [I64, I32, I32])
_sig("gcnb_head_f32", I32, [P, P, P, I64, I32, I32, U32, I32, P, P, P, P, P, I64, P])
_sig("gcnb_head_tc_supported", I32, [I32, I32])
_sig("gcnb_head_tc_f32", I32, [P, P, P, I64, I32, I32, U32, I32, P, P, P, P, P, I64, P])
_sig("gcnb_head_reduce_dw_f32", I32, [P, P, I64, I32, I32, P])
_sig("gcnb_adam_step_f32", I32, [P, F32, F32, F32, F32, F32, P])
_sig("gcnb_sumsq_workspace", I64, [I64])
_sig("gcnb_sumsq_f32", I32, [P, I64, P, P, P])


E_BADARG, E_UNSUPPORTED = 10001, 10002  # GCNB_E_* of include/gcnb.h


class GcnbError(RuntimeError):
    pass


def check(code):
    if code != 0:
        raise GcnbError("libgcn_b200: error %d: %s" % (code, lib.gcnb_error_string(code).decode()))


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def device_check():
    n = C.c_int(0)
    check(lib.gcnb_device_check(C.byref(n)))
    return n.value


def make_rng(seed, history=(), elem_offset=0):
    """history: iterable of (n_elements_of_an_earlier_rng_op, times_it_ran); elem_offset: global index of local element 0."""
    r = RngT()
    r.seed = int(seed) & 0xFFFFFFFF
    r.group_offset, r.elem_lead = elem_offset // 4, elem_offset % 4
    merged = {}
    for size, count in history:
        g = (int(size) + 3) // 4
        merged[g] = merged.get(g, 0) + int(count)
    items = [(g, c) for g, c in merged.items() if c > 0]
    if len(items) > MAX_RNG_HIST:
        raise ValueError("too many distinct RNG consumers")
    r.n_hist = len(items)
    for i, (g, c) in enumerate(items):
        r.hist_groups[i] = g
        r.hist_count[i] = c
    return r


class SpmmPlan:
    """gcnb_spmm_plan: load-balancing metadata of one CSR index (borrowed device arrays)."""

    def __init__(self, indptr, indices, n_cols, seg_nnz=0):
        self.indptr, self.indices = indptr, indices  # keep alive: the plan borrows them
        self.n_rows = indptr.numel() - 1
        self.n_cols = int(n_cols)
        h = C.c_void_p()
        check(lib.gcnb_spmm_plan_create(ptr(indptr), ptr(indices), self.n_rows, self.n_cols, int(seg_nnz), stream(),
                                        C.byref(h)))
        self.h = h

    def info(self):
        out = (I64 * 8)()
        check(lib.gcnb_spmm_plan_info(self.h, out))
        keys = ("n_rows", "nnz", "n_seg", "n_split_rows", "n_slots", "n_queues", "seg_nnz", "max_deg")
        return dict(zip(keys, [int(x) for x in out]))

    def stage(self, values, dim, h_indptr=None, h_indices=None, window_rows=0, min_seg=0, seg_cap=0, min_window_nnz=0):
        """window-staged fast path for a static value array (GraphSum); h_*: optional numpy uint32 host copies."""
        self._staged_values = values  # keep alive: the staged path is keyed on this pointer
        hp = h_indptr.ctypes.data_as(C.c_void_p) if h_indptr is not None else None
        hi = h_indices.ctypes.data_as(C.c_void_p) if h_indices is not None else None
        check(lib.gcnb_spmm_plan_stage_ex(self.h, hp, hi, ptr(values), int(dim), window_rows, min_seg, seg_cap,
                                          min_window_nnz, stream()))
        return self.stage_info()

    def stage_async_begin(self, values, dim=16):
        """background staging (helper thread + own stream); returns a job handle (None: nothing to do)"""
        self._staged_values = values
        job = C.c_void_p()
        check(lib.gcnb_spmm_plan_stage_async_begin(self.h, ptr(values), int(dim), C.byref(job)))
        return job if job.value else None

    def stage_async_finish(self, job):
        check(lib.gcnb_spmm_plan_stage_async_finish(self.h, job))
        return self.stage_info()

    def stage_info(self):
        out = (I64 * 8)()
        check(lib.gcnb_spmm_plan_stage_info(self.h, out))
        keys = ("staged", "window_rows", "staged_nnz", "rem_nnz", "n_segs", "n_runs", "n_blocks", "n_slots")
        return dict(zip(keys, [int(x) for x in out]))

    def spmm(self, values, B, C_out, dim, perm=None):
        check(lib.gcnb_spmm_f32(self.h, ptr(values), ptr(perm), ptr(B), ptr(C_out), int(dim), stream()))
        return C_out

    def attach_bittile(self, bt, values):
        """route 16-column products with this value tensor through a BitTilePlan (None detaches)"""
        self._bt, self._bt_values = bt, values  # keep both alive
        check(lib.gcnb_spmm_plan_attach_bittile(self.h, bt.h if bt is not None else None, ptr(values)))

    def set_own_cols(self, col0, col1):
        """row-partitioned product: columns [col0, col1) are this rank's own slab of B (call before stage())"""
        check(lib.gcnb_spmm_plan_set_own_cols(self.h, int(col0), int(col1)))

    def stage_own(self, values, B_own, dim=16):
        """launch the staged runs whose windows lie inside the own slab, from that slab; True if anything was launched"""
        launched = C.c_int(0)
        check(lib.gcnb_spmm_stage_own_f32(self.h, ptr(values), ptr(B_own), int(dim), stream(), C.byref(launched)))
        return bool(launched.value)

    def spmm_ld(self, values, B, ldb, C_out, ldc, dim, perm=None, b_off=0, c_off=0):
        """product on the column slab [off, off + dim) of wider row-major matrices (row strides ldb / ldc floats)"""
        check(lib.gcnb_spmm_ld_f32(self.h, ptr(values), ptr(perm), C.c_void_p(B.data_ptr() + 4 * int(b_off)), int(ldb),
                                   C.c_void_p(C_out.data_ptr() + 4 * int(c_off)), int(ldc), int(dim), stream()))
        return C_out

    def close(self):
        if getattr(self, "h", None) and lib is not None:
            lib.gcnb_spmm_plan_destroy(self.h)
            self.h = None

    __del__ = close


def stage_host_build(indptr, indices, n_cols, dim=16, window_rows=0, min_seg=0, seg_cap=0, min_window_nnz=0, n_cta=0,
                     n_threads=0, own_cols=(0, 0)):
    """Host-only run of the staging builder (no CUDA): returns the plan arrays as numpy (see csrc/spmm_plan.cuh)."""
    import numpy as np
    indptr = np.ascontiguousarray(indptr, np.uint32)
    indices = np.ascontiguousarray(indices, np.uint32)
    h = C.c_void_p()
    check(lib.gcnb_stage_host_build_own(indptr.ctypes.data_as(C.c_void_p), indices.ctypes.data_as(C.c_void_p),
                                        len(indptr) - 1, int(n_cols), dim, window_rows, min_seg, seg_cap, min_window_nnz,
                                        n_cta, n_threads, int(own_cols[0]), int(own_cols[1]), C.byref(h)))
    try:
        sz = (I64 * 13)()
        check(lib.gcnb_stage_host_sizes(h, sz))
        keys = ("window_rows", "n_win", "n_cta", "staged_nnz", "n_bundles", "n_runs", "n_blocks", "n_slots", "rem_nnz",
                "n_rows", "nnz", "n_segs", "n_own_runs")
        out = dict(zip(keys, [int(x) for x in sz]))
        n_rows = out["n_rows"]

        def grab(which, n, dtype):
            a = np.zeros(n, dtype)
            if n:
                check(lib.gcnb_stage_host_copy(h, which, a.ctypes.data_as(C.c_void_p), a.nbytes))
            return a
        out["bundles"] = grab(0, out["n_bundles"] * 4, np.uint32).reshape(-1, 4)
        out["runs"] = grab(1, out["n_runs"] * 4, np.uint32).reshape(-1, 4)
        out["run_begin"] = grab(2, out["n_cta"] + 1, np.uint32)
        out["pidx"] = grab(3, out["n_blocks"] * 128, np.uint16)
        out["pperm"] = grab(4, out["n_blocks"] * 128, np.uint32)
        out["row_slot"] = grab(5, n_rows + 1, np.uint32)
        out["r_indptr"] = grab(6, n_rows + 1, np.uint32)
        out["r_indices"] = grab(7, out["rem_nnz"], np.uint32)
        out["r_perm"] = grab(8, out["rem_nnz"], np.uint32)
        out["lens"] = grab(9, out["n_bundles"] * 32, np.uint16)
        out["lane_slot"] = grab(10, out["n_bundles"] * 32, np.uint32)
        out["own_runs"] = grab(11, out["n_own_runs"] * 4, np.uint32).reshape(-1, 4)
        out["own_run_begin"] = grab(12, out["n_cta"] + 1, np.uint32)
        return out
    finally:
        lib.gcnb_stage_host_destroy(h)


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def bittile_host_build(indptr, indices, values, n_cols, row_scale=None, col_scale=None, min_tile_nnz=0, n_cta=0, n_threads=0,
                       chunk_cols=0, row_blocks=0):
    """Host-only run of the bit-tile builder (no CUDA): plan arrays as numpy (BitTileHost, csrc/spmm_bittile.cu)."""
    import numpy as np
    indptr = np.ascontiguousarray(indptr, np.uint32)
    indices = np.ascontiguousarray(indices, np.uint32)
    values = np.ascontiguousarray(values, np.float32)
    rs = None if row_scale is None else np.ascontiguousarray(row_scale, np.float32)
    cs = None if col_scale is None else np.ascontiguousarray(col_scale, np.float32)
    h = C.c_void_p()
    check(lib.gcnb_bittile_host_build(_np_ptr(indptr), _np_ptr(indices), _np_ptr(values), len(indptr) - 1, int(n_cols),
                                      _np_ptr(rs), _np_ptr(cs), min_tile_nnz, chunk_cols, row_blocks, n_cta, n_threads, C.byref(h)))
    try:
        sz = (I64 * 12)()
        check(lib.gcnb_bittile_host_sizes(h, sz))
        keys = ("n_rows", "n_cols", "nnz", "n_blk", "n_tiles", "tile_nnz", "n_items", "n_cta", "chunk", "rb", "n_unfactored",
                "rem_nnz")
        out = dict(zip(keys, [int(x) for x in sz]))

        def grab(which, n, dtype):
            a = np.zeros(n, dtype)
            if n:
                check(lib.gcnb_bittile_host_copy(h, which, a.ctypes.data_as(C.c_void_p), a.nbytes))
            return a
        out["tile_chunk"] = grab(0, out["n_tiles"], np.uint32)
        wpr, rb = out["chunk"] // 64, out["rb"]
        out["bits"] = grab(1, out["n_tiles"] * rb * 128 * wpr, np.uint64).reshape(-1, rb, 128, wpr)
        out["cta_tile_ptr"] = grab(2, out["n_cta"] + 1, np.uint32)
        out["cta_item_ptr"] = grab(3, out["n_cta"] + 1, np.uint32)
        out["items"] = grab(4, out["n_items"] * 2, np.uint32).reshape(-1, 2)
        out["r_indptr"] = grab(5, out["n_rows"] + 1, np.uint32)
        out["rem_nnz"] = int(out["r_indptr"][-1]) if out["n_rows"] else 0
        out["r_indices"] = grab(6, out["rem_nnz"], np.uint32)
        out["r_values"] = grab(7, out["rem_nnz"], np.float32)
        out["row_scale"] = grab(8, out["n_rows"], np.float32)
        out["col_scale"] = grab(9, out["n_cols"], np.float32)
        return out
    finally:
        lib.gcnb_bittile_host_destroy(h)


def ell_host_build(indptr, indices, n_cols, n_threads=0):
    """Host-only run of the pattern-only ELL builder (EllHost, csrc/spmm_ell.cu): plan arrays as numpy."""
    import numpy as np
    indptr = np.ascontiguousarray(indptr, np.uint32)
    indices = np.ascontiguousarray(indices, np.uint32)
    h = P()
    check(lib.gcnb_ell_host_build(_np_ptr(indptr), _np_ptr(indices), len(indptr) - 1, int(n_cols), n_threads, C.byref(h)))
    try:
        sz = (I64 * 8)()
        check(lib.gcnb_ell_host_sizes(h, sz))
        keys = ("n_rows", "n_cols", "nnz", "n_bundles", "idx_words", "n_split", "n_slots", "wide_min")
        out = dict(zip(keys, [int(x) for x in sz]))

        def grab(which, count):
            a = np.zeros(count, np.uint32)
            if count:
                check(lib.gcnb_ell_host_copy(h, which, a.ctypes.data_as(C.c_void_p), a.nbytes))
            return a
        out["idx"] = grab(0, out["idx_words"])
        out["off"] = grab(1, out["n_bundles"] + 1)
        out["steps"] = grab(2, out["n_bundles"])
        out["rows"] = grab(3, out["n_bundles"] * 8)
        out["split_row"] = grab(4, out["n_split"])
        out["split_ptr"] = grab(5, out["n_split"] + 1)
        return out
    finally:
        lib.gcnb_ell_host_destroy(h)


class EllPlan:
    """gcnb_ell_plan: R = diag(row_scale) * pattern * B2 at width 16 (host numpy CSR pattern in, device plan)."""

    def __init__(self, indptr, indices, n_cols):
        import numpy as np
        indptr = np.ascontiguousarray(indptr, np.uint32)
        indices = np.ascontiguousarray(indices, np.uint32)
        self.n_rows, self.n_cols = len(indptr) - 1, int(n_cols)
        h = P()
        check(lib.gcnb_ell_plan_create(_np_ptr(indptr), _np_ptr(indices), self.n_rows, self.n_cols, stream(), C.byref(h)))
        self.h = h

    def gather16(self, B2, row_scale, R_out):
        """B2: (n_cols + 1) x 16 device tensor whose last row is zero"""
        check(lib.gcnb_ell_gather16_f32(self.h, ptr(B2), ptr(row_scale), ptr(R_out), stream()))

    def close(self):
        if getattr(self, "h", None):
            lib.gcnb_ell_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


class BitTilePlan:
    """gcnb_bittile_plan: tensor-core GraphSum on 128 x 64 bit-map tiles + remainder CSR (host numpy CSR in, device plan)."""

    def __init__(self, indptr, indices, values, n_cols, row_scale=None, col_scale=None, min_tile_nnz=0, chunk_cols=0,
                 row_blocks=0):
        import numpy as np
        indptr = np.ascontiguousarray(indptr, np.uint32)
        indices = np.ascontiguousarray(indices, np.uint32)
        values = None if values is None else np.ascontiguousarray(values, np.float32)  # None: a pattern scaled by rs x cs
        rs = None if row_scale is None else np.ascontiguousarray(row_scale, np.float32)
        cs = None if col_scale is None else np.ascontiguousarray(col_scale, np.float32)
        self.h = C.c_void_p()
        check(lib.gcnb_bittile_plan_create(_np_ptr(indptr), _np_ptr(indices), _np_ptr(values), len(indptr) - 1, int(n_cols),
                                           _np_ptr(rs), _np_ptr(cs), min_tile_nnz, chunk_cols, row_blocks, stream(), C.byref(self.h)))

    @classmethod
    def from_device(cls, d_indptr, d_indices, d_values, n_rows, n_cols, d_row_scale=None, d_col_scale=None, min_tile_nnz=0,
                    chunk_cols=0, row_blocks=0):
        """gcnb_bittile_plan_create_device: the same plan built on the GPU from device tensors (None when the matrix needs
        the host builder, GCNB_E_UNSUPPORTED)"""
        self = cls.__new__(cls)
        self.h = C.c_void_p()
        rc = lib.gcnb_bittile_plan_create_device(ptr(d_indptr), ptr(d_indices), ptr(d_values), int(n_rows), int(n_cols),
                                                 ptr(d_row_scale), ptr(d_col_scale), min_tile_nnz, chunk_cols, row_blocks,
                                                 stream(), C.byref(self.h))
        if rc == E_UNSUPPORTED:
            self.h = None
            return None
        check(rc)
        return self

    ARRAYS = ("tile_chunk", "bits", "cta_tile_ptr", "cta_item_ptr", "items", "row_scale", "col_scale", "ell_idx", "ell_off",
              "ell_steps", "ell_rows", "ell_split_row", "ell_split_ptr")

    def arrays(self):
        """every device array of the plan as numpy (test aid: the host-built and the device-built plan must be identical)"""
        import numpy as np
        sz = (I64 * 16)()
        check(lib.gcnb_bittile_plan_sizes(self.h, sz))
        out = {"ell_slots": int(sz[13]), "tile_nnz": int(sz[14]), "rem_nnz": int(sz[15])}
        for which, name in enumerate(self.ARRAYS):
            dt = np.uint64 if which in (1, 4) else np.uint32  # scales compared as bit patterns
            a = np.zeros(int(sz[which]), dt)
            if a.size:
                check(lib.gcnb_bittile_plan_copy(self.h, which, a.ctypes.data_as(C.c_void_p), a.nbytes))
            out[name] = a
        return out

    def info(self):
        out = (I64 * 8)()
        check(lib.gcnb_bittile_plan_info(self.h, out))
        keys = ("n_tiles", "tile_nnz", "rem_nnz", "n_blk", "chunk", "n_cta", "bitmap_bytes", "packed_bytes")
        d = dict(zip(keys, [int(x) for x in out]))
        d["ell"], d["rb"], d["chunk"] = d["chunk"] // 100000, (d["chunk"] // 1000) % 100, d["chunk"] % 1000
        return d

    def spmm16(self, B, C_out):
        check(lib.gcnb_bittile_spmm16_f32(self.h, ptr(B), ptr(C_out), stream()))

    def spmm_ld(self, B, ldb, C_out, ldc, dim, b_off=0, c_off=0):
        """product on the column slab [off, off + dim) of wider row-major matrices, 16 columns at a time"""
        check(lib.gcnb_bittile_spmm_ld_f32(self.h, C.c_void_p(B.data_ptr() + 4 * int(b_off)), int(ldb),
                                           C.c_void_p(C_out.data_ptr() + 4 * int(c_off)), int(ldc), int(dim), stream()))

    def debug_parts(self, mask):
        check(lib.gcnb_bittile_debug_parts(self.h, int(mask)))

    def debug_pack(self, B):
        import numpy as np
        out = np.zeros(self.info()["packed_bytes"] // 2, np.uint16)
        check(lib.gcnb_bittile_debug_pack(self.h, ptr(B), out.ctypes.data_as(C.c_void_p), out.nbytes, stream()))
        return out

    def close(self):
        if getattr(self, "h", None) and lib is not None:
            lib.gcnb_bittile_plan_destroy(self.h)
            self.h = None

    __del__ = close


class Csc:
    """gcnb_csc: transposed view (column pointers, row ids, permutation into the CSR value array)."""

    def __init__(self, indptr, indices, n_cols):
        import torch
        self.n_rows, self.n_cols = indptr.numel() - 1, int(n_cols)
        h = C.c_void_p()
        check(lib.gcnb_csc_create(ptr(indptr), ptr(indices), self.n_rows, self.n_cols, stream(), C.byref(h)))
        self.h = h
        cp, ri, pm, dense = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int(0)
        check(lib.gcnb_csc_arrays(h, C.byref(cp), C.byref(ri), C.byref(pm), C.byref(dense)))
        self.is_dense = bool(dense.value)
        self.colptr_ptr, self.rowidx_ptr, self.perm_ptr = cp, ri, pm
        self.nnz = int(indices.numel())
        self.plan = None
        if not self.is_dense:
            # wrap the library-owned arrays as torch tensors without copying (for plan creation)
            self.colptr = _wrap_u32(cp.value, self.n_cols + 1, indptr.device)
            self.rowidx = _wrap_u32(ri.value, max(self.nnz, 1), indptr.device)[: self.nnz]
            self.perm = _wrap_u32(pm.value, max(self.nnz, 1), indptr.device)[: self.nnz]
            self.plan = SpmmPlan(self.colptr, self.rowidx, self.n_rows)

    def close(self):
        if self.plan is not None:
            self.plan.close()
            self.plan = None
        if getattr(self, "h", None) and lib is not None:
            lib.gcnb_csc_destroy(self.h)
            self.h = None

    __del__ = close


def _wrap_u32(addr, n, device):
    """torch view (int32 storage) over library-owned device memory."""
    import torch

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (addr, False), "version": 2}
    return torch.as_tensor(h, device=device)


def matmul_nn(A, B, Cout, m, n, p):
    check(lib.gcnb_matmul_nn_f32(ptr(A), ptr(B), ptr(Cout), m, n, p, stream()))
    return Cout


def matmul_nt(dC, B, dA, m, n, p):
    check(lib.gcnb_matmul_nt_f32(ptr(dC), ptr(B), ptr(dA), m, n, p, stream()))
    return dA


def matmul_tn(A, dC, dB, m, n, p, ws=None):
    import torch
    need = lib.gcnb_matmul_tn_workspace(m, n, p)
    if ws is None or ws.numel() * ws.element_size() < need:
        ws = torch.empty((need + 3) // 4, dtype=torch.float32, device=A.device)
    check(lib.gcnb_matmul_tn_f32(ptr(A), ptr(dC), ptr(dB), m, n, p, ptr(ws), ws.numel() * ws.element_size(), stream()))
    return dB


def dropout_maskbits(bits, n_rows, f, p, rng):
    check(lib.gcnb_dropout_maskbits(ptr(bits), n_rows, f, p, C.byref(rng), stream()))


def dense_feat_fwd(X, bits, p_drop, W, out, n, f, p):
    check(lib.gcnb_dense_feat_fwd_f32(ptr(X), ptr(bits), p_drop, ptr(W), ptr(out), n, f, p, stream()))


def dense_feat_tn(X, bits, p_drop, dH, dW, n, f, p, ws=None):
    import torch
    need = lib.gcnb_dense_feat_tn_workspace(n, f, p)
    if ws is None or ws.numel() * ws.element_size() < need:
        ws = torch.empty((need + 3) // 4, dtype=torch.float32, device=X.device)
    check(lib.gcnb_dense_feat_tn_f32(ptr(X), ptr(bits), p_drop, ptr(dH), ptr(dW), n, f, p, ptr(ws),
                                     ws.numel() * ws.element_size(), stream()))


def dense_tc_pack_x(X, n, f):
    """packed bf16 x 3 operand image of X for gcnb_dense_tc_fwd_f32 (once per dataset)"""
    import torch
    img = torch.empty(lib.gcnb_dense_tc_x_bytes(n, f), dtype=torch.uint8, device=X.device)
    check(lib.gcnb_dense_tc_pack_x(ptr(X), ptr(img), n, f, stream()))
    return img


def dense_tc_fwd(x_img, W, out, n, f, p, ws=None):
    import torch
    need = lib.gcnb_dense_tc_w_bytes(f, p)
    if ws is None:
        ws = torch.empty(need, dtype=torch.uint8, device=W.device)
    check(lib.gcnb_dense_tc_fwd_f32(ptr(x_img), ptr(W), ptr(out), n, f, p, ptr(ws), ws.numel(), stream()))
    return out


def dense_tc_pack_xt(X, n, f):
    import torch
    img = torch.empty(lib.gcnb_dense_tc_xt_bytes(n, f), dtype=torch.uint8, device=X.device)
    check(lib.gcnb_dense_tc_pack_xt(ptr(X), ptr(img), n, f, stream()))
    return img


def dense_tc_tn(xt_img, dH, dW, n, f, p, ws=None):
    import torch
    if ws is None:
        ws = torch.empty(lib.gcnb_dense_tc_tn_workspace(n, f, p), dtype=torch.uint8, device=dH.device)
    check(lib.gcnb_dense_tc_tn_f32(ptr(xt_img), ptr(dH), ptr(dW), n, f, p, ptr(ws), ws.numel(), stream()))
    return dW


def glorot(w, rows, cols, rng):
    check(lib.gcnb_glorot_f32(ptr(w), w.numel(), rows, cols, C.byref(rng), stream()))
    return w


def dropout_fwd(x, mask, p, rng=None, ext_mask=None):
    check(lib.gcnb_dropout_fwd_f32(ptr(x), ptr(mask), ptr(ext_mask), x.numel(), p,
                                   C.byref(rng) if rng is not None else None, stream()))


def dropout_bwd(g, mask, p):
    check(lib.gcnb_dropout_bwd_f32(ptr(g), ptr(mask), g.numel(), p, stream()))


def relu_fwd(x, mask, training):
    check(lib.gcnb_relu_fwd_f32(ptr(x), ptr(mask), x.numel(), int(training), stream()))


def relu_bwd(g, mask):
    check(lib.gcnb_relu_bwd_f32(ptr(g), ptr(mask), g.numel(), stream()))


def relu_dropout_fwd(x, mask, p, training, rng=None, ext_mask=None):
    check(lib.gcnb_relu_dropout_fwd_f32(ptr(x), ptr(mask), ptr(ext_mask), x.numel(), p, int(training),
                                        C.byref(rng) if rng is not None else None, stream()))


def relu_dropout_bwd(g, mask, p):
    check(lib.gcnb_relu_dropout_bwd_f32(ptr(g), ptr(mask), g.numel(), p, stream()))


def set_truth(truth, split, label, cur):
    check(lib.gcnb_set_truth(ptr(truth), ptr(split), ptr(label), truth.numel(), cur, stream()))


def zeroed_workspace(nbytes, device):
    import torch
    return torch.zeros((int(nbytes) + 3) // 4, dtype=torch.int32, device=device)


def softmax_ce(logits, grad, truth, n, num_classes, num_samples, training, result, ws):
    check(lib.gcnb_softmax_ce_f32(ptr(logits), ptr(grad), ptr(truth), n, num_classes, num_samples, int(training),
                                  ptr(result), ptr(ws), stream()))


def head(y, w, truth, n, in_dim, num_classes, num_samples, training, logits, grad, dy, dw, result, ws, tensor_cores=False):
    """Output head in one kernel (csrc/head.cu): logits = y W, softmax cross-entropy + counts, dy = dz W^T, dW = y^T dz."""
    fn = lib.gcnb_head_tc_f32 if tensor_cores else lib.gcnb_head_f32
    check(fn(ptr(y), ptr(w), ptr(truth), n, in_dim, num_classes, num_samples, int(training), ptr(logits),
                            ptr(grad), ptr(dy), ptr(result), ptr(ws), ws.numel() * 4, stream()))
    if training:
        check(lib.gcnb_head_reduce_dw_f32(ptr(ws), ptr(dw), n, in_dim, num_classes, stream()))


def adam_step(tensors, weight_decay, beta1, beta2, eps, step_size):
    """tensors: list of (w, g, m, v, decay)."""
    t = AdamTensorsT()
    t.n_tensors = len(tensors)
    for i, (w, g, m, v, decay) in enumerate(tensors):
        t.w[i], t.g[i], t.m[i], t.v[i] = w.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
        t.size[i] = w.numel()
        t.decay[i] = int(bool(decay))
    check(lib.gcnb_adam_step_f32(C.byref(t), weight_decay, beta1, beta2, eps, step_size, stream()))


def sumsq(w, out, ws):
    check(lib.gcnb_sumsq_f32(ptr(w), w.numel(), ptr(out), ptr(ws), stream()))
