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


def exact_ground_truth(X, Q, k):
    """fp64 exact kNN ids by L2 with (dist, id) ties -- numpy, for recall only."""
    X64 = X.astype(np.float64)
    out = np.empty((Q.shape[0], k), np.int64)
    xn = (X64 * X64).sum(1)
    for s in range(0, Q.shape[0], 256):
        q = Q[s:s + 256].astype(np.float64)
        d = xn[None, :] - 2.0 * q @ X64.T + (q * q).sum(1)[:, None]
        idx = np.argpartition(d, k, axis=1)[:, :k]
        dd = np.take_along_axis(d, idx, 1)
        order = np.lexsort((idx, dd), axis=1)
        out[s:s + 256] = np.take_along_axis(idx, order, 1)
    return out
