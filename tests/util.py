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
