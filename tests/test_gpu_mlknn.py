"""GPU: knn_classify / knn_regress (NeuronDB/src/ml/ml_knn.c:112-357, 363-569) through the C ABI.

The reference computes euclidean_distance (:76-90: float difference, double sum of squares in dimension order,
sqrt) of every (feature, label) row to the query, qsorts by distance and votes among / averages the k nearest.  The
restatement below does exactly that in numpy; the GPU path (scan kernel in the same arithmetic + top-k, vote on the
host) must agree on every query whose k-th and (k+1)-th distances are not a near tie (qsort's order of equal
distances is unspecified in the reference itself)."""
import numpy as np
import pytest

import oracle_lib as O
import workloads as W

pytestmark = pytest.mark.gpu


def reference_neighbours(X, Q, k):
    order, clear = [], []
    for q in Q:
        d = (q[None, :] - X).astype(np.float64)          # a[i] - b[i] is a float operation, then widened
        acc = np.zeros(len(X))
        for j in range(X.shape[1]):
            acc += d[:, j] * d[:, j]
        dist = np.sqrt(acc)
        o = np.argsort(dist, kind="stable")
        order.append(o[:k])
        clear.append(len(X) == k or dist[o[k]] - dist[o[k - 1]] > 1e-6 * dist[o[k]])
    return np.array(order), np.array(clear)


@pytest.mark.parametrize("n,dim,nq,k", [(5000, 16, 200, 5), (1200, 96, 64, 1), (300, 7, 50, 25), (40, 3, 10, 40)])
def test_knn_classify_and_regress_equal_the_reference_loops(ndb, n, dim, nq, k):
    rng = np.random.default_rng(n + k)
    X = W.mixture(n, dim, 6, 10 + n)
    Q = W.mixture(nq, dim, 6, 11 + n, centers_seed=10 + n)
    labels = rng.integers(0, 2, n).astype(np.float64)
    labels[rng.integers(0, n, max(1, n // 50))] = 7.0       # labels outside {0, 1} take no part in the vote
    targets = rng.standard_normal(n) * 100.0
    ds = ndb.Dataset(dim)
    ds.append(X)
    got_c = ds.knn_classify(labels, Q, k)
    got_r = ds.knn_regress(targets, Q, k)
    nb, clear = reference_neighbours(X, Q, k)
    assert clear.mean() > 0.9
    want_c, want_r = [], []
    for o in nb:
        votes = [0.0, 0.0]
        for i in o:
            c = int(labels[i])
            if 0 <= c < 2:
                votes[c] += 1.0
        want_c.append(1 if votes[1] > votes[0] else 0)
        s = 0.0
        for i in o:
            s += targets[i]
        want_r.append(s / k)
    want_c, want_r = np.array(want_c), np.array(want_r)
    assert np.array_equal(got_c[clear], want_c[clear])
    assert np.array_equal(got_r[clear], want_r[clear])           # same neighbours in the same order: the same double


def test_knn_classify_errors_are_the_sql_functions(ndb):
    X = W.gaussian(20, 4, 1)
    ds = ndb.Dataset(4)
    ds.append(X)
    lab = np.zeros(20)
    with pytest.raises(ndb.NdbError) as e:
        ds.knn_classify(lab, X[:2], 0)                   # "k must be at least 1"
    assert e.value.code == -1 and "k must be at least 1" in str(e.value)
    with pytest.raises(ndb.NdbError) as e:
        ds.knn_regress(lab, X[:2], 21)                   # "need at least 21 samples, got 20"
    assert e.value.code == -8 and "need at least 21 samples" in str(e.value)
    bad = X[:2].copy()
    bad[1, 2] = np.nan
    with pytest.raises(ndb.NdbError) as e:
        ds.knn_classify(lab, bad, 3)
    assert e.value.code == -4
    lab[:] = 1.0
    assert ds.knn_classify(lab, X[:3], 4).tolist() == [1, 1, 1]
    lab[:] = [i % 2 for i in range(20)]
    tie = ds.knn_classify(lab, X[:5], 20)                # 10 : 10 -> class 0 (strict majority for 1)
    assert tie.tolist() == [0] * 5


def test_knn_classify_and_regress_equal_the_reference_outputs(ndb):
    """tests/golden/ml_paths.npz: classes and means the reference's OWN KNNSample / euclidean_distance / compare_samples /
    qsort produced (oracle/_ref/libndb_ref_leafs.so, written by tests/golden/make_golden.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ml_paths.npz"))
    for n, dim, k, seed in [(400, 8, 5, 3), (900, 33, 4, 7), (1500, 16, 12, 5), (60, 3, 6, 60)]:
        X = W.mixture(n, dim, max(2, k // 2), seed)
        Q = W.mixture(30, dim, 5, seed + 1, centers_seed=seed)
        lab = np.random.default_rng(seed).integers(0, 3, n).astype(np.float64)
        ds = ndb.Dataset(dim)
        ds.append(X)
        assert np.array_equal(ds.knn_classify(lab, Q, min(k, n)), g["knn_cls_n%d" % n])
        assert np.array_equal(ds.knn_regress(lab, Q, min(k, n)), g["knn_mean_n%d" % n])


def test_knn_order_is_the_order_of_the_doubles(ndb):
    """Rows whose distances are distinct doubles but the same float: the reference sorts the doubles (KNNSample.distance),
    so the neighbours are not the lowest row ids of the float tie."""
    n_tied = 40
    X = np.zeros((100, 2), np.float32)
    X[:, 0] = 50.0 + np.arange(100)                                  # far rows
    X[:n_tied, 0] = 1.0
    X[:n_tied, 1] = (n_tied - np.arange(n_tied)) * 1e-6              # sqrt(1 + e^2): 1.0f for all of them, e decreasing with the row
    q = np.zeros((1, 2), np.float32)
    ds = ndb.Dataset(2)
    ds.append(X, ids=np.arange(100, dtype=np.int64)[::-1] * 1000 + 7)   # labels follow the row position, not the id
    targets = np.arange(100, dtype=np.float64)
    assert ds.knn_regress(targets, q, 5)[0] == (39 + 38 + 37 + 36 + 35) / 5
    lab = (np.arange(100) >= 37).astype(np.float64)                  # rows 37, 38, 39 vote 1, rows 35, 36 vote 0
    assert ds.knn_classify(lab, q, 5)[0] == 1
    cls, mean, rows = O.knn_ml(X, targets, q, 5)
    assert rows[0].tolist() == [39, 38, 37, 36, 35]
