"""GPU: cluster_kmeans (NeuronDB/src/ml/ml_kmeans.c:146-303, kmeanspp_init :45-139) through the C ABI.

Labels, centres (bit patterns), the seeds kmeanspp_init picked and the number of Lloyd iterations must equal (a) the
committed outputs of the reference's OWN functions (tests/golden/ml_paths.npz, written from oracle/_ref/
libndb_ref_leafs.so by tests/golden/make_golden.py) and (b) the oracle restatement on larger seeded inputs."""
import os

import numpy as np
import pytest

import oracle_lib as O
import workloads as W

pytestmark = pytest.mark.gpu
BITS = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ml_paths.npz")
CASES = [(400, 8, 5, 3), (900, 33, 4, 7), (1500, 16, 12, 5), (60, 3, 6, 60)]          # test_oracle._ml_cases


def test_cluster_kmeans_equals_the_reference_outputs(ndb):
    g = np.load(GOLDEN)
    for n, dim, k, seed in CASES:
        tag = "n%d" % n
        X = W.mixture(n, dim, max(2, k // 2), seed)
        labels, centers, seeds, it = ndb.cluster_kmeans(X, k, 0, g["draws_" + tag])
        assert it == int(g["km_iters_" + tag]), tag
        assert np.array_equal(labels, g["km_labels_" + tag]), tag
        assert np.array_equal(BITS(centers), g["km_center_bits_" + tag]), tag


@pytest.mark.parametrize("n,dim,k,max_iters", [(20000, 32, 16, 0), (5000, 7, 64, 4), (3000, 130, 9, 2), (4097, 5, 2, 0), (64, 4, 64, 3)])
def test_cluster_kmeans_equals_the_oracle(ndb, n, dim, k, max_iters):
    X = W.mixture(n, dim, max(2, k // 3), n + k)
    if n == 64:
        X[5:9] = X[4]                          # duplicate rows: zero weights in the walk, clusters that end empty
    draws = np.random.default_rng(n).integers(0, O.RAND_MAX, k, dtype=np.int64).astype(np.int32)
    if n == 4097:
        draws[1] = O.RAND_MAX                  # r = sum: the walk runs to the last rows (or past them: the :103-113 branch)
    if n == 5000:
        draws[2] = 0                           # r = 0: the first unselected row
    want_l, want_c, want_s, want_it = O.cluster_kmeans(X, k, max_iters, draws)
    labels, centers, seeds, it = ndb.cluster_kmeans(X, k, max_iters, draws)
    assert np.array_equal(seeds, want_s)
    assert it == want_it and np.array_equal(labels, want_l)
    assert np.array_equal(BITS(centers), BITS(want_c))
    assert labels.min() >= 1 and labels.max() <= k


def test_cluster_kmeans_errors_are_the_sql_functions(ndb):
    X = W.gaussian(10, 4, 1)
    d = np.arange(12, dtype=np.int32)
    with pytest.raises(ndb.NdbError) as e:
        ndb.cluster_kmeans(X, 1, 5, d[:1])
    assert e.value.code == -1 and "number of clusters must be at least 2" in str(e.value)
    with pytest.raises(ndb.NdbError) as e:
        ndb.cluster_kmeans(X, 11, 5, d[:11])
    assert e.value.code == -1 and "not enough vectors for cluster count (need >= 11, have 10)" in str(e.value)
    bad = X.copy()
    bad[3, 1] = np.inf
    with pytest.raises(ndb.NdbError) as e:
        ndb.cluster_kmeans(bad, 2, 5, d[:2])
    assert e.value.code == -4


def test_cluster_kmeans_certified_pick_equals_the_literal_walk(ndb):
    """The D^2-weighted draw is normally read off a parallel prefix sum with a certificate; NDB_CKM_SEQUENTIAL forces the
    reference's loop (one thread, row order) for every seed.  Same seeds, labels and centres either way, and the default
    path needs the walk only where the certificate cannot hold (r = 0: the draw of 0 below)."""
    n, dim, k = 30000, 12, 24
    X = W.mixture(n, dim, 8, 99)
    draws = np.random.default_rng(99).integers(0, O.RAND_MAX, k, dtype=np.int64).astype(np.int32)
    draws[5] = 0
    want = O.cluster_kmeans(X, k, 3, draws)
    got = ndb.cluster_kmeans(X, k, 3, draws)
    walked = ndb.last_kernel_stats()[2]
    os.environ["NDB_CKM_SEQUENTIAL"] = "1"
    try:
        lit = ndb.cluster_kmeans(X, k, 3, draws)
        walked_lit = ndb.last_kernel_stats()[2]
    finally:
        del os.environ["NDB_CKM_SEQUENTIAL"]
    for a, b, c in zip(want, got, lit):
        assert np.array_equal(np.asarray(a), np.asarray(b)) and np.array_equal(np.asarray(a), np.asarray(c))
    assert walked_lit == k - 1 and 1 <= walked <= 3


# ---- cluster_minibatch_kmeans (ml_minibatch_kmeans.c:67-198, 206-449) ---------------------------------------
MB_CASES = [(800, 8, 5, 50, 20, 21), (1500, 24, 12, 100, 30, 22), (300, 6, 7, 1000, 5, 23), (40, 3, 8, 16, 12, 24)]   # test_oracle._minibatch_cases


def _mb_rows(n, dim, k, seed):
    X = W.mixture(n, dim, max(2, k // 2), seed)
    if n == 40:
        X[:] = X[:4].repeat(10, axis=0)          # 4 distinct rows for 8 clusters: the seeding stops early, 4 centroids stay zero
    return X


def test_cluster_minibatch_kmeans_equals_the_reference_outputs(ndb):
    """Labels and centre bits the reference's OWN minibatch_kmeans_pp_init + main-loop text produced (golden ml_paths.npz),
    and the number of rand() calls it made."""
    g = np.load(GOLDEN)
    for n, dim, k, batch, iters, seed in MB_CASES:
        X = _mb_rows(n, dim, k, seed)
        labels, centers, used = ndb.cluster_minibatch_kmeans(X, k, batch, iters, g["mb_draws_n%d" % n])
        assert np.array_equal(labels, g["mb_labels_n%d" % n]), n
        assert np.array_equal(BITS(centers), g["mb_center_bits_n%d" % n]), n
        assert used == (k if n != 40 else 4) + min(batch, n) * iters


@pytest.mark.parametrize("n,dim,k,batch,iters", [(30000, 32, 16, 100, 100), (5000, 130, 40, 256, 7), (2000, 5, 3, 1, 50)])
def test_cluster_minibatch_kmeans_equals_the_oracle(ndb, n, dim, k, batch, iters):
    X = W.mixture(n, dim, max(2, k // 3), n + k)
    draws = np.random.default_rng(n).integers(0, O.RAND_MAX, k + batch * iters, dtype=np.int64).astype(np.int32)
    want_l, want_c, want_used = O.cluster_minibatch_kmeans(X, k, batch, iters, draws)
    labels, centers, used = ndb.cluster_minibatch_kmeans(X, k, batch, iters, draws)
    assert used == want_used and np.array_equal(labels, want_l) and np.array_equal(BITS(centers), BITS(want_c))


def test_cluster_minibatch_kmeans_errors_are_the_sql_functions(ndb):
    X = W.gaussian(10, 4, 1)
    d = np.arange(64, dtype=np.int32)
    for k, batch, msg in ((1, 5, "num_clusters must be at least 2"), (3, 0, "batch_size must be at least 1"),
                          (11, 5, "Not enough vectors (10) for 11 clusters")):
        with pytest.raises(ndb.NdbError) as e:
            ndb.cluster_minibatch_kmeans(X, k, batch, 2, d)
        assert e.value.code == -1 and msg in str(e.value)
    with pytest.raises(ndb.NdbError) as e:
        ndb.cluster_minibatch_kmeans(X, 2, 5, 2, d[:3])          # the draws run out: the callback reports it
    assert e.value.code == -1
