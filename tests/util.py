import numpy as np


def to_dev(a, dev):
    """numpy -> torch on device; uint32 arrays travel as int32 bit patterns."""
    import torch
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    return torch.from_numpy(a).to(dev)


def to_np(t, dtype=None):
    a = t.detach().cpu().numpy()
    return a if dtype is None else a.view(dtype)


def assert_close(got, want, rtol=1e-5, atol=None, what=""):
    """relative 1e-5 (north_star) with an absolute floor tied to the data scale for cancellation-prone sums."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if atol is None:
        atol = 1e-6 * (np.abs(want).max() if want.size else 1.0) + 1e-30
    err = np.abs(got - want)
    tol = atol + rtol * np.abs(want)
    bad = err > tol
    if bad.any():
        i = np.argmax(err - tol)
        raise AssertionError("%s: %d/%d elements off; worst idx %d got %r want %r (err %.3e tol %.3e)" %
                             (what, bad.sum(), bad.size, i, got.flat[i], want.flat[i], err.flat[i], tol.flat[i]))


def random_csr(rng, n_rows, n_cols, mean_deg, heavy_rows=(), empty_rows=()):
    deg = rng.poisson(mean_deg, n_rows).astype(np.int64)
    for r, d in heavy_rows:
        deg[r] = d
    for r in empty_rows:
        deg[r] = 0
    indptr = np.zeros(n_rows + 1, np.uint32)
    indptr[1:] = np.cumsum(deg)
    indices = rng.integers(0, n_cols, int(indptr[-1])).astype(np.uint32)
    return indptr, indices


def pubmed_root(root):
    """pubmed ships without its .svmlight (reference .gitignore): a synthetic one in SURVEY 8d's shape (500 dims, 3 classes,
    30-70 nnz per row, row-normalised values, seed 20230606) next to links to the shipped .graph / .split.  Returns a
    directory laid out like the repository root (data/pubmed.*)."""
    import os
    import tempfile
    out = tempfile.mkdtemp(prefix="gcnb_pubmed_")
    os.makedirs(os.path.join(out, "data"))
    for ext in ("graph", "split"):
        os.symlink(os.path.join(root, "data", "pubmed." + ext), os.path.join(out, "data", "pubmed." + ext))
    rng = np.random.default_rng(20230606)
    n = sum(1 for _ in open(os.path.join(root, "data", "pubmed.graph")))
    with open(os.path.join(out, "data", "pubmed.svmlight"), "w") as f:
        for _ in range(n):
            k = int(rng.integers(30, 71))
            cols = np.sort(rng.choice(500, k, replace=False))
            f.write("%d %s\n" % (rng.integers(0, 3), " ".join("%d:%.6f" % (c, 1.0 / k) for c in cols)))
    return out
