"""CPU study for the bit-tile GraphSum: error of the dense-block product when B' = s_j * B_j is carried as three
bf16 pieces (exact), two fp16 pieces under one power-of-two scale per launch, or two bf16 pieces, against the parity bar
1e-5 * |ref| + 1e-6 * max|ref|.  1/8-scale bench graph, exact (float64) accumulation of the quantised operand.
r1 result: bf16x3 and fp16x2 stay below 4 % of the bar (also for operands spanning e^+-4 in row magnitude), bf16x2 exceeds it."""
import sys, numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as ge
ge.load_package()
from parallel_gcn_b200 import engine as eng
n, m = 232965 // 8, 57307946 // 8
indptr, indices = eng.synth_graph(n, m, n_blocks=6)
values = eng.synth_graph_values(indptr, indices, 0, np.diff(indptr).astype(np.uint32))
deg = np.diff(indptr.astype(np.int64)); rows = np.repeat(np.arange(n), deg)
s = (1.0 / np.sqrt(deg.astype(np.float32))).astype(np.float32)
rng = np.random.default_rng(0)
for name, x in (("N(0,1) activations", rng.standard_normal((n, 16)).astype(np.float32)),
                ("wide dynamic range (gradients)", (rng.standard_normal((n, 16)) * np.exp(rng.normal(0, 4, (n, 1)))).astype(np.float32) * 1e-6)):
    ref = np.zeros((n, 16)); np.add.at(ref, rows, values[:, None].astype(np.float64) * x[indices])
    Bp = (s[:, None] * x).astype(np.float32)
    def spmm(Bq):  # exact sums of the quantised operand, then the row scale
        out = np.zeros((n, 16)); np.add.at(out, rows, Bq[indices]); return s[:, None].astype(np.float64) * out
    # bf16 x 3 (exact)
    exact = spmm(Bp.astype(np.float64))
    # fp16 x 2 with a power-of-two scale
    k = 14 - int(np.ceil(np.log2(np.abs(Bp).max())))
    S = np.float32(2.0 ** k)
    hi = (Bp * S).astype(np.float16); lo = (Bp * S - hi.astype(np.float32)).astype(np.float16)
    q2 = (hi.astype(np.float64) + lo.astype(np.float64)) / float(S)
    f16 = spmm(q2)
    # bf16 x 2 (hi truncated, mid rounded)
    hb = (Bp.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32); r1 = Bp - hb
    mb = ((r1.view(np.uint32) + np.uint32(0x8000)) & np.uint32(0xFFFF0000)).view(np.float32)
    b2 = spmm((hb.astype(np.float64) + mb))
    scale = np.abs(ref).max()
    tol = 1e-5 * np.abs(ref) + 1e-6 * scale
    for tag, got in (("bf16x3", exact), ("fp16x2", f16), ("bf16x2", b2)):
        err = np.abs(got - ref)
        print("%-32s %-7s max err/scale %.2e  worst err/tol %.3f  elements over tol %d" % (name, tag, err.max() / scale, (err / tol).max(), int((err > tol).sum())))
