"""ctypes front-end of the parity oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product (parallel-gcn_b200/) never does.

  * `lib`  -> oracle/liboracle.so   : gcn_oracle.c, the plain-C restatement (reference file:line cited there)
  * `ref`  -> oracle/_ref/libref_cpu.so : the reference's own CPU code compiled in place (may be absent)
  * `OracleGCN` : the GCN driver loop (GCN::train_epoch / eval / run) restated on top of the C primitives,
                  in two flavours: "ref_cpu" (hpdga-spring23/src/gcn.cpp:64-250) and "ref_gpu" (src/gcn.cu:146-455).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
f32, i32, u32, u8, i64 = np.float32, np.int32, np.uint32, np.uint8, np.int64


def build():
    subprocess.check_call(["make", "-s", "-C", HERE], stdout=subprocess.DEVNULL)


def _load(path):
    return C.CDLL(path) if os.path.exists(path) else None


if not os.path.exists(os.path.join(HERE, "liboracle.so")):
    build()
lib = _load(os.path.join(HERE, "liboracle.so"))
ref = _load(os.path.join(HERE, "_ref", "libref_cpu.so"))

P = C.c_void_p


def _p(a):
    return None if a is None else a.ctypes.data_as(P)


def _sig(l, name, res, args):
    fn = getattr(l, name)
    fn.restype = res
    fn.argtypes = args


_sig(lib, "orc_xorshift_set", None, [C.c_uint64, C.c_uint64])
_sig(lib, "orc_xorshift_get", None, [P])
_sig(lib, "orc_xorshift_next", C.c_uint32, [])
_sig(lib, "orc_xorshift_seed_from_libc", None, [])
_sig(lib, "orc_libc_srand", None, [C.c_uint])
_sig(lib, "orc_philox4x32_10", None, [P, P, P])
_sig(lib, "orc_curand_uniform4", None, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, P])
_sig(lib, "orc_glorot_xorshift", None, [C.c_int64, C.c_int, C.c_int, P])
_sig(lib, "orc_glorot_philox", None, [C.c_int64, C.c_uint, C.c_uint, C.c_uint32, P, C.c_int, P])
_sig(lib, "orc_dataset_parse", P, [C.c_char_p])
_sig(lib, "orc_dataset_dims", None, [P, P])
_sig(lib, "orc_dataset_copy", None, [P, C.c_int, P])
_sig(lib, "orc_dataset_free", None, [P])
_sig(lib, "orc_graph_values", None, [C.c_int64, P, P, P])
_sig(lib, "orc_graphsum", None, [C.c_int64, C.c_int, P, P, P, P, P])
_sig(lib, "orc_spmm", None, [C.c_int64, C.c_int, P, P, P, P, P])
_sig(lib, "orc_spmm_bwd", None, [C.c_int64, C.c_int64, C.c_int, P, P, P, P, P])
_sig(lib, "orc_matmul", None, [C.c_int64, C.c_int, C.c_int, P, P, P])
_sig(lib, "orc_matmul_bwd", None, [C.c_int64, C.c_int, C.c_int, P, P, P, P, P])
_sig(lib, "orc_relu_fwd", None, [C.c_int64, P, P, C.c_int])
_sig(lib, "orc_relu_bwd", None, [C.c_int64, P, P])
_sig(lib, "orc_dropout_mask_xorshift", None, [C.c_int64, C.c_float, P])
_sig(lib, "orc_dropout_mask_philox", None, [C.c_int64, C.c_float, C.c_uint32, P, C.c_int, P])
_sig(lib, "orc_dropout_apply", None, [C.c_int64, P, P, C.c_float])
_sig(lib, "orc_dropout_scale", C.c_float, [C.c_float, C.c_int])
_sig(lib, "orc_cross_entropy", C.c_float, [C.c_int64, C.c_int, P, P, P, C.c_int64, C.c_int, P])
_sig(lib, "orc_wrong_count", C.c_int64, [C.c_int64, C.c_int, P, P, P])
_sig(lib, "orc_sumsq", C.c_float, [C.c_int64, P])
_sig(lib, "orc_set_truth", None, [C.c_int64, P, P, C.c_uint32, P])
_sig(lib, "orc_adam_step_size", C.c_float, [C.c_float, C.c_float, C.c_float, C.c_int])
_sig(lib, "orc_adam_step", None, [C.c_int64, P, P, P, P, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float])
_sig(lib, "orc_fnv1a64", C.c_uint64, [P, C.c_int64])

if ref is not None:
    _sig(ref, "ref_srand", None, [C.c_uint])
    _sig(ref, "ref_rand_state_get", None, [P])
    _sig(ref, "ref_rand_state_set", None, [C.c_uint64, C.c_uint64])
    _sig(ref, "ref_rand_next", C.c_uint32, [])
    _sig(ref, "ref_dataset_parse", P, [C.c_char_p, C.c_char_p])
    _sig(ref, "ref_dataset_from_arrays", P, [C.c_int64, C.c_int64, P, P, C.c_int64, P, P, P, P, P, C.c_int, C.c_int])
    _sig(ref, "ref_dataset_dims", None, [P, P])
    _sig(ref, "ref_dataset_copy", None, [P, C.c_int, P])
    _sig(ref, "ref_dataset_free", None, [P])
    _sig(ref, "ref_gcn_create", P, [P, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int])
    _sig(ref, "ref_gcn_free", None, [P])
    _sig(ref, "ref_gcn_train_epoch", None, [P, P])
    _sig(ref, "ref_gcn_eval", None, [P, C.c_int, P])
    _sig(ref, "ref_gcn_num_variables", C.c_int, [P])
    _sig(ref, "ref_gcn_variable_size", C.c_int64, [P, C.c_int, C.c_int])
    _sig(ref, "ref_gcn_variable_get", None, [P, C.c_int, C.c_int, P])
    _sig(ref, "ref_gcn_variable_set", None, [P, C.c_int, P])
    _sig(ref, "ref_gcn_run", C.c_double, [P, P])
    _sig(ref, "ref_timer_total", C.c_float, [C.c_int])
    _sig(ref, "ref_timer_reset", None, [])
    _sig(ref, "ref_graphsum", None, [C.c_int64, C.c_int, P, P, P, P, P, P])
    _sig(ref, "ref_sparse_matmul", None, [C.c_int64, C.c_int, C.c_int, P, P, P, P, P, P, P])
    _sig(ref, "ref_matmul", None, [C.c_int64, C.c_int, C.c_int, P, P, P, P, P, P])
    _sig(ref, "ref_cross_entropy", C.c_float, [C.c_int64, C.c_int, P, P, P, C.c_int])
    _sig(ref, "ref_dropout", None, [C.c_int64, C.c_float, P, P])
    _sig(ref, "ref_relu", None, [C.c_int64, P, P])
    _sig(ref, "ref_glorot", None, [C.c_int64, C.c_int, C.c_int, P])
    _sig(ref, "ref_adam", None, [C.c_int64, C.c_int, P, P, C.c_int, C.c_float, C.c_float])


def fnv(a):
    a = np.ascontiguousarray(a)
    return "%016x" % lib.orc_fnv1a64(_p(a), a.nbytes)


class Dataset:
    """Host CSR of one dataset as the reference parser builds it (uint32 indices, int32 labels)."""

    FIELDS = ("g_indptr", "g_indices", "f_indptr", "f_indices", "f_value", "label", "split")

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def num_nodes(self):
        return len(self.g_indptr) - 1

    def split_counts(self):
        return tuple(int((self.split == s).sum()) for s in (1, 2, 3))

    def graph_values(self):
        out = np.empty(len(self.g_indices), f32)
        lib.orc_graph_values(self.num_nodes, _p(self.g_indptr), _p(self.g_indices), _p(out))
        return out


def parse_dataset(prefix):
    """oracle parser (gcn_oracle.c orc_dataset_parse); prefix = path without extension."""
    h = lib.orc_dataset_parse(prefix.encode())
    if not h:
        return None
    dims = np.zeros(10, i64)
    lib.orc_dataset_dims(h, _p(dims))
    n, gnnz, frows, fnnz, in_dim, out_dim, nsplit = (int(x) for x in dims[:7])
    arrs = [np.empty(n + 1, u32), np.empty(gnnz, u32), np.empty(frows + 1, u32), np.empty(fnnz, u32),
            np.empty(fnnz, f32), np.empty(frows, i32), np.empty(nsplit, u32)]
    for k, a in enumerate(arrs):
        lib.orc_dataset_copy(h, k, _p(a))
    lib.orc_dataset_free(h)
    return Dataset(**dict(zip(Dataset.FIELDS, arrs)), input_dim=in_dim, output_dim=out_dim)


def ref_parse_dataset(root_dir, name):
    """the reference's own Parser (hpdga-spring23/src/parser.cpp) via oracle/_ref."""
    h = ref.ref_dataset_parse(root_dir.encode(), name.encode())
    if not h:
        return None, None
    dims = np.zeros(8, i64)
    ref.ref_dataset_dims(h, _p(dims))
    n, gnnz, frows, fnnz, in_dim, out_dim, nsplit, nlabel = (int(x) for x in dims)
    arrs = [np.empty(n + 1, i32), np.empty(gnnz, i32), np.empty(frows + 1, i32), np.empty(fnnz, i32),
            np.empty(fnnz, f32), np.empty(nlabel, i32), np.empty(nsplit, i32)]
    for k, a in enumerate(arrs):
        ref.ref_dataset_copy(h, k, _p(a))
    return h, Dataset(**dict(zip(Dataset.FIELDS, arrs)), input_dim=in_dim, output_dim=out_dim)


def ref_dataset_from(ds):
    g = [np.ascontiguousarray(getattr(ds, k)).astype(i32) for k in ("g_indptr", "g_indices", "f_indptr", "f_indices")]
    fv = np.ascontiguousarray(ds.f_value, f32)
    lab = np.ascontiguousarray(ds.label).astype(i32)
    sp = np.ascontiguousarray(ds.split).astype(i32)
    return ref.ref_dataset_from_arrays(ds.num_nodes, len(g[1]), _p(g[0]), _p(g[1]), len(g[3]), _p(g[2]), _p(g[3]),
                                       _p(fv), _p(lab), _p(sp), ds.input_dim, ds.output_dim)


class PhiloxStream:
    """Bookkeeping of cuRAND's per-state draw counters: every RNG op of `size` elements
    advances states 0..ceil(size/4)-1 by one draw (src/variable.cu:49, src/module.cu:25)."""

    def __init__(self, seed, max_size):
        self.seed = int(seed)
        self.draws = np.zeros((int(max_size) + 3) // 4, u32)

    def consume(self, size):
        g = (int(size) + 3) // 4
        cur = self.draws[:g].copy()
        self.draws[:g] += 1
        return cur


class OracleGCN:
    """L-layer GCN training loop restated from the reference drivers.

    flavour "ref_cpu": hpdga-spring23/src/gcn.cpp (L=2 only there; xorshift RNG seeded from libc rand();
                       CE normalised by counted labelled rows; coef recomputed per edge).
    flavour "ref_gpu": src/gcn.cu (any L; Philox streams; CE normalised by split counts; graph_value hoisted).
    Weights / masks can be injected (`weights=`, `mask_fn=`) so both engines see the same randomness.
    """

    def __init__(self, ds, hidden_dims=(16,), dropouts=(0.5, 0.5), lr=0.01, weight_decay=5e-4, beta1=0.9,
                 beta2=0.999, eps=1e-8, flavour="ref_gpu", seed=19990304, weights=None, mask_fn=None,
                 libc_seed=None, fused_uniform=True):
        self.ds, self.flavour = ds, flavour
        self.N, self.F, self.Cn = ds.num_nodes, ds.input_dim, ds.output_dim
        self.dims = [self.F] + [int(h) for h in hidden_dims] + [self.Cn]
        self.L = len(self.dims) - 1
        assert len(dropouts) == self.L
        self.dropouts = [f32(p) for p in dropouts]
        self.lr, self.wd, self.b1, self.b2, self.eps = f32(lr), f32(weight_decay), f32(beta1), f32(beta2), f32(eps)
        self.mask_fn, self.fused = mask_fn, int(fused_uniform)
        self.gpu = flavour == "ref_gpu"
        self.values = ds.graph_values() if self.gpu else None
        self.counts = ds.split_counts()
        self.step = 0
        fnnz = len(ds.f_indices)
        if self.gpu:
            rand_sizes = [fnnz] + [self.dims[l] * self.dims[l + 1] for l in range(self.L)] + \
                         [self.N * self.dims[l + 1] for l in range(self.L - 1)]
            self.philox = PhiloxStream(seed, max(rand_sizes))
        else:
            if libc_seed is not None:
                lib.orc_libc_srand(libc_seed)
            lib.orc_xorshift_seed_from_libc()  # gcn.cpp:65 init_rand_state()
        self.W, self.m, self.v = [], [], []
        for l in range(self.L):
            r, c = self.dims[l], self.dims[l + 1]
            w = np.empty(r * c, f32)
            if weights is not None:
                w[:] = np.asarray(weights[l], f32).ravel()
            elif self.gpu:
                lib.orc_glorot_philox(w.size, r, c, self.philox.seed, _p(self.philox.consume(w.size)), self.fused, _p(w))
            else:
                lib.orc_glorot_xorshift(w.size, r, c, _p(w))
            self.W.append(w)
            self.m.append(np.zeros_like(w))
            self.v.append(np.zeros_like(w))
        self.truth = np.empty(self.N, i32)
        self.trace = {}

    # -- pieces -----------------------------------------------------------------------------------
    def _dropout(self, x, p, want_mask, tag):
        if self.mask_fn is not None:
            mask = np.ascontiguousarray(self.mask_fn(tag, x.size, float(p)), u8)
        else:
            mask = np.empty(x.size, u8)
            if self.gpu:
                lib.orc_dropout_mask_philox(x.size, p, self.philox.seed, _p(self.philox.consume(x.size)), self.fused, _p(mask))
            else:
                lib.orc_dropout_mask_xorshift(x.size, p, _p(mask))
        scale = lib.orc_dropout_scale(p, int(self.gpu))
        lib.orc_dropout_apply(x.size, _p(x), _p(mask), scale)
        return (mask, scale) if want_mask else (None, scale)

    def _graphsum(self, x, dim):
        out = np.empty(self.N * dim, f32)
        lib.orc_graphsum(self.N, dim, _p(self.ds.g_indptr), _p(self.ds.g_indices), _p(self.values), _p(x), _p(out))
        return out

    def forward(self, split, training):
        ds = self.ds
        lib.orc_set_truth(self.N, _p(ds.split), _p(ds.label), split, _p(self.truth))
        st = {"split": split}
        x = ds.f_value.copy()  # set_input
        if training:
            st["mask_in"], st["scale_in"] = self._dropout(x, self.dropouts[0], False, "input")
        st["x"] = x
        h = np.empty(self.N * self.dims[1], f32)
        lib.orc_spmm(self.N, self.dims[1], _p(ds.f_indptr), _p(ds.f_indices), _p(x), _p(self.W[0]), _p(h))
        st["var1_0"] = h
        acts = []
        for l in range(self.L):
            d = self.dims[l + 1]
            if l > 0:
                hin = acts[-1]
                h = np.empty(self.N * d, f32)
                lib.orc_matmul(self.N, self.dims[l], d, _p(hin), _p(self.W[l]), _p(h))
                st["var1_%d" % l] = h
            z = self._graphsum(h, d)
            if l < self.L - 1:
                rmask = np.zeros(z.size, u8)
                lib.orc_relu_fwd(z.size, _p(z), _p(rmask), int(training))
                st["relu_%d" % l] = rmask
                if training:
                    st["dmask_%d" % l], st["dscale_%d" % l] = self._dropout(z, self.dropouts[l + 1], True, "hidden%d" % l)
                acts.append(z)
                st["var2_%d" % l] = z
        logits = z
        grad = np.empty_like(logits) if training else None
        ns = {1: self.counts[0], 2: self.counts[1], 3: self.counts[2]}[split] if self.gpu else 0
        cnt = np.zeros(1, i64)
        loss = f32(lib.orc_cross_entropy(self.N, self.Cn, _p(logits), _p(self.truth), _p(grad), ns, int(training), _p(cnt)))
        tot = np.zeros(1, i64)
        wrong = lib.orc_wrong_count(self.N, self.Cn, _p(logits), _p(self.truth), _p(tot))
        l2 = f32(lib.orc_sumsq(self.W[0].size, _p(self.W[0])))
        if self.gpu:  # GCN::finalize src/gcn.cu:440-455
            total = ns
            loss = f32(loss / f32(total)) + f32(self.wd * l2 / f32(2))
            # ref-GPU counts `wrong` with the shifted-logit test over labelled rows only
            acc = f32(f32(np.uint32(total - wrong)) / f32(total))
        else:  # gcn.cpp:188-189
            total = int(tot[0])
            loss = f32(loss + f32(self.wd * l2 / f32(2)))
            acc = f32(f32(total - wrong) / f32(total))
        st.update(logits=logits, grad=grad, acts=acts)
        self.trace = st
        return float(loss), float(acc)

    def backward_and_step(self):
        st, ds = self.trace, self.ds
        g = st["grad"]
        wgrads = [None] * self.L
        for l in range(self.L - 1, -1, -1):
            d = self.dims[l + 1]
            g1 = self._graphsum(g, d)  # var1.grad = A * var2.grad
            if l > 0:
                a = st["acts"][l - 1]
                ga = np.empty(self.N * self.dims[l], f32)
                gw = np.empty(self.dims[l] * d, f32)
                lib.orc_matmul_bwd(self.N, self.dims[l], d, _p(a), _p(self.W[l]), _p(g1), _p(ga), _p(gw))
                wgrads[l] = gw
                lib.orc_dropout_apply(ga.size, _p(ga), _p(st["dmask_%d" % (l - 1)]), st["dscale_%d" % (l - 1)])
                lib.orc_relu_bwd(ga.size, _p(ga), _p(st["relu_%d" % (l - 1)]))
                g = ga
            else:
                gw = np.empty(self.F * d, f32)
                lib.orc_spmm_bwd(self.N, self.F, d, _p(ds.f_indptr), _p(ds.f_indices), _p(st["x"]), _p(g1), _p(gw))
                wgrads[0] = gw
        self.wgrads = wgrads
        self.step += 1
        ss = lib.orc_adam_step_size(self.lr, self.b1, self.b2, self.step)
        for l in range(self.L):
            lib.orc_adam_step(self.W[l].size, _p(self.W[l]), _p(wgrads[l]), _p(self.m[l]), _p(self.v[l]), int(l == 0),
                              self.wd, self.b1, self.b2, self.eps, ss)

    def train_epoch(self):
        r = self.forward(1, True)
        self.backward_and_step()
        return r

    def eval(self, split):
        return self.forward(split, False)
