"""Seeded synthetic inputs of the shapes BASELINE.json names, scaled by argument."""
import numpy as np


def gaussian(n, dim, seed):
    return np.random.default_rng(seed).standard_normal((n, dim), dtype=np.float32)


def mixture(n, dim, components, seed, sigma=0.3, centers_seed=None):
    """n draws from a `components`-component Gaussian mixture (centres N(0,1), within sigma)."""
    crng = np.random.default_rng(seed if centers_seed is None else centers_seed)
    centers = crng.standard_normal((components, dim), dtype=np.float32)
    rng = np.random.default_rng(seed + 1000003)
    which = rng.integers(0, components, size=n)
    X = centers[which] + sigma * rng.standard_normal((n, dim), dtype=np.float32)
    return np.ascontiguousarray(X, dtype=np.float32)


def normalised(n, dim, seed):
    X = gaussian(n, dim, seed)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    return np.ascontiguousarray(X, dtype=np.float32)


def exact_ground_truth(X, Q, k, metric=1):
    """fp64 exact kNN ids (metric 1 = L2, 3 = inner product: largest dot first) with (dist, id) ties --
    numpy, for recall only.  Rows are processed in chunks so that 10 M-row workloads fit in memory."""
    nq = Q.shape[0]
    q = Q.astype(np.float64)
    best_d = np.full((nq, k), np.inf)
    best_i = np.full((nq, k), -1, np.int64)
    chunk = 1_000_000
    for r0 in range(0, X.shape[0], chunk):
        X64 = X[r0:r0 + chunk].astype(np.float64)
        if metric == 3:
            d = -(q @ X64.T)
        else:
            d = (X64 * X64).sum(1)[None, :] - 2.0 * (q @ X64.T) + (q * q).sum(1)[:, None]
        kk = min(k, d.shape[1])
        idx = np.argpartition(d, kk - 1, axis=1)[:, :kk]
        cd = np.concatenate([best_d, np.take_along_axis(d, idx, 1)], axis=1)
        ci = np.concatenate([best_i, idx + r0], axis=1)
        order = np.lexsort((ci, cd), axis=1)[:, :k]
        best_d = np.take_along_axis(cd, order, 1)
        best_i = np.take_along_axis(ci, order, 1)
    return best_i
