"""GPU parity for the IVF path (ivfbuild / ivfinsert / ivfgettuple) against the oracle."""
import numpy as np
import pytest

import workloads as W

pytestmark = pytest.mark.gpu

BITS = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)


def build_pair(ndb, orc, X, lists, metric=1, ids=None):
    ix = ndb.IvfIndex(X.shape[1], lists, metric)
    ix.ivfbuild(X)
    ns = orc.lib().orc_ivf_train_samples(X.shape[0], lists)
    oC, _, _, _, _ = orc.kmeans_train(X[:ns], lists)
    assert np.array_equal(BITS(ix.centroids()), BITS(oC)), "k-means centroids differ from the oracle"
    got_lists = ix.ivfinsert(X, ids)
    want_lists = orc.ivf_assign(X, oC)
    assert np.array_equal(got_lists, want_lists), "ivfinsert list assignment differs"
    off, rows = orc.lists_from_assignment(want_lists, lists)
    assert np.array_equal(ix.list_sizes(), np.diff(off))
    return ix, oC, off, rows


@pytest.mark.parametrize("n,dim,lists,nprobe,k", [(20000, 32, 64, 8, 10), (30000, 128, 100, 10, 10),
                                                   (5000, 17, 16, 16, 5), (8000, 96, 50, 3, 32)])
@pytest.mark.parametrize("metric", [1, 2, 3])
def test_ivf_full_scan_parity(ndb, orc, n, dim, lists, nprobe, k, metric):
    X = W.mixture(n, dim, lists, 2024 + n)
    Q = W.mixture(300, dim, lists, 2025 + n, centers_seed=2024 + n)
    ix, oC, off, rows = build_pair(ndb, orc, X, lists, metric)
    probes = ix.select_clusters(Q, nprobe)
    want_probes = np.stack([orc.select_clusters(q, oC, nprobe) for q in Q[:50]])
    assert np.array_equal(probes[:50, :want_probes.shape[1]], want_probes)
    d, i = ix.search(Q, nprobe, k, ndb.IVF_FULL)
    od, oi, _ = orc.ivf_search(X, oC, off, rows, Q, nprobe, k, strategy=metric, literal=False)
    assert np.array_equal(i, oi)
    assert np.array_equal(BITS(d), BITS(od))


def test_ivf_literal_mode_parity(ndb, orc):
    """ivfCollectCandidates as written: k*10 candidate cap in probe order."""
    X = W.mixture(20000, 64, 40, 7)
    Q = W.mixture(200, 64, 40, 8, centers_seed=7)
    ix, oC, off, rows = build_pair(ndb, orc, X, 40)
    d, i = ix.search(Q, 10, 10, ndb.IVF_LITERAL)
    od, oi, _ = orc.ivf_search(X, oC, off, rows, Q, 10, 10, literal=True)
    assert np.array_equal(i, oi) and np.array_equal(BITS(d), BITS(od))


def test_ivf_ids_and_unsorted_insert(ndb, orc):
    """heap ids that are not the row index, inserted in two batches in shuffled id order."""
    X = W.mixture(6000, 24, 20, 31)
    ids = np.random.default_rng(1).permutation(100000)[:6000].astype(np.int64)
    Q = W.mixture(100, 24, 20, 32, centers_seed=31)
    ix = ndb.IvfIndex(24, 20)
    ix.ivfbuild(X)
    ix.ivfinsert(X[:2500], ids[:2500])
    ix.ivfinsert(X[2500:], ids[2500:])
    assert len(ix) == 6000
    oC = ix.centroids()
    assign = orc.ivf_assign(X, oC)
    off, rows = orc.lists_from_assignment(assign, 20)
    d, i = ix.search(Q, 5, 10)
    od, oi, _ = orc.ivf_search(X, oC, off, rows, Q, 5, 10, ids=ids)
    assert np.array_equal(i, oi) and np.array_equal(BITS(d), BITS(od))
    # literal mode depends on insertion order, not id order
    d, i = ix.search(Q, 5, 10, ndb.IVF_LITERAL)
    od, oi, _ = orc.ivf_search(X, oC, off, rows, Q, 5, 10, ids=ids, literal=True)
    assert np.array_equal(i, oi) and np.array_equal(BITS(d), BITS(od))


def test_ivf_edge_cases(ndb, orc):
    X = W.gaussian(400, 8, 3)
    ix = ndb.IvfIndex(8, 4)
    with pytest.raises(ndb.NdbError):          # search before build: no centroids
        ix.search(X[:2], 2, 5)
    with pytest.raises(ndb.NdbError):          # "ivf: not enough sample vectors" (ivf_am.c:596-603)
        ndb.IvfIndex(8, 500).ivfbuild(X)
    ix.ivfbuild(X)
    d, i = ix.search(X[:3], 2, 5)              # trained but empty: no rows
    assert np.all(i == -1) and np.all(np.isinf(d))
    ix.ivfinsert(X[:3])                        # fewer rows than k
    d, i = ix.search(X[:3], 4, 5)
    assert np.all(i[:, 0] == np.arange(3)) and np.all(d[:, 0] == 0.0)
    assert np.all(i[:, 3:] == -1)
    d, i = ix.search(X[:3], 99, 5)             # nprobe > nlists is clamped (ivf_am.c:1622-1623)
    assert np.all(i[:, 3:] == -1)


def test_ivf_config2_shape_recall(ndb, orc):
    """BASELINE config 2 shape scaled to 200k rows: recall@10 vs exact ground truth, id parity
    with the oracle on a query slice."""
    n, dim, lists, nprobe = 200_000, 128, 256, 16
    X = W.mixture(n, dim, lists, 2024)
    Q = W.mixture(2000, dim, lists, 2025, centers_seed=2024)
    ix = ndb.IvfIndex(dim, lists)
    ix.ivfbuild(X)
    ix.ivfinsert(X)
    d, i = ix.search(Q, nprobe, 10)
    gt = W.exact_ground_truth(X, Q, 10)
    recall = orc.recall_at_k(i, gt)
    oC = ix.centroids()
    off, rows = orc.lists_from_assignment(orc.ivf_assign(X, oC), lists)
    od, oi, _ = orc.ivf_search(X, oC, off, rows, Q[:200], nprobe, 10)
    assert np.array_equal(i[:200], oi) and np.array_equal(BITS(d[:200]), BITS(od))
    assert recall >= 0.95, recall
    df, i_f = ix.search(Q, nprobe, 10, arith=ndb.ARITH_FAST)
    assert orc.recall_at_k(i_f, gt) >= recall - 0.002
    assert np.max(np.abs(df - d) / np.maximum(d, 1e-6)) < 1e-5


def test_pipelined_search_equals_synchronous(ndb):
    """ndb_b200_ivf_search_begin / _end: two batches in flight return what the synchronous call returns."""
    X = W.mixture(20000, 32, 64, 91)
    ix = ndb.IvfIndex(32, 64)
    ix.ivfbuild(X)
    ix.ivfinsert(X)
    batches = [np.ascontiguousarray(W.mixture(300, 32, 64, 92 + b, centers_seed=91)) for b in range(5)]
    want = [ix.search(Q, 8, 10) for Q in batches]
    for arith in (ndb.ARITH_IVF_F32, ndb.ARITH_TENSOR):
        ref = want if arith == ndb.ARITH_IVF_F32 else [ix.search(Q, 8, 10, ndb.IVF_FULL, arith) for Q in batches]
        outs = [(np.empty((300, 10), np.float32), np.empty((300, 10), np.int64)) for _ in batches]
        prev = None
        for b, Q in enumerate(batches):
            t = ix.search_begin(Q, outs[b][0], outs[b][1], 8, 10, ndb.IVF_FULL, arith)
            if prev is not None:
                ix.search_end(prev)
            prev = t
        ix.search_end(prev)
        for (d, i), (wd, wi) in zip(outs, ref):
            assert np.array_equal(i, wi) and np.array_equal(d.view(np.uint32), wd.view(np.uint32))
    # a third batch without an end is refused; a NaN query is reported by end
    o = [(np.empty((300, 10), np.float32), np.empty((300, 10), np.int64)) for _ in range(3)]
    t0 = ix.search_begin(batches[0], o[0][0], o[0][1], 8, 10)
    t1 = ix.search_begin(batches[1], o[1][0], o[1][1], 8, 10)
    with pytest.raises(ndb.NdbError):
        ix.search_begin(batches[2], o[2][0], o[2][1], 8, 10)
    ix.search_end(t0)
    ix.search_end(t1)
    bad = batches[0].copy()
    bad[7, 3] = np.nan
    t = ix.search_begin(bad, o[0][0], o[0][1], 8, 10)
    with pytest.raises(ndb.NdbError):
        ix.search_end(t)


def test_ivf_literal_mode_at_the_largest_k(ndb, orc):
    """k = 128: the literal kernel's candidate arrays (12 * 4 * k * 10 = 61 440 bytes) exceed the 48 KB default of
    dynamic shared memory, so the launch needs the opt-in attribute."""
    X = W.mixture(30000, 32, 16, 17)
    Q = W.mixture(40, 32, 16, 18, centers_seed=17)
    ix, oC, off, rows = build_pair(ndb, orc, X, 16)
    d, i = ix.search(Q, 8, 128, ndb.IVF_LITERAL)
    od, oi, _ = orc.ivf_search(X, oC, off, rows, Q, 8, 128, literal=True)
    assert np.array_equal(i, oi) and np.array_equal(BITS(d), BITS(od))
    with pytest.raises(ndb.NdbError):          # the literal mode is ivfCollectCandidates' own arithmetic only
        ix.search(Q, 8, 10, ndb.IVF_LITERAL, ndb.ARITH_FAST)


def test_wrappers_report_dimension_errors(ndb):
    """A mismatched array is the reference's 'dimensions must match' error (EDIM), not an out-of-bounds read."""
    X = W.gaussian(600, 8, 3)
    ix = ndb.IvfIndex(8, 4)
    ix.ivfbuild(X)
    ix.ivfinsert(X)
    for call in (lambda: ix.search(W.gaussian(3, 9, 1), 2, 5), lambda: ix.ivfinsert(W.gaussian(3, 7, 1)),
                 lambda: ix.select_clusters(W.gaussian(3, 16, 1), 2), lambda: ix.assign(W.gaussian(3, 4, 1))):
        with pytest.raises(ndb.NdbError) as e:
            call()
        assert e.value.code == -5
    ds = ndb.Dataset(8)
    ds.append(X)
    with pytest.raises(ndb.NdbError) as e:
        ds.knn(W.gaussian(2, 5, 1), 3)
    assert e.value.code == -5


@pytest.mark.parametrize("n,dim,lists", [(60000, 128, 1024), (30000, 96, 512), (20000, 40, 300), (5000, 260, 256)])
def test_list_assignment_on_the_tensor_cores_equals_the_oracle(ndb, orc, n, dim, lists):
    """ivfinsert's nearest-centroid loop (ivf_am.c:906-935) for centroid sets large enough to go through the tensor
    cores (GEMM form, certified per row): the list of EVERY row equals the oracle's sequential fp32 loop -- strict <,
    lowest index -- on ordinary (not bf16-representable) rows, including rows that coincide with a centroid."""
    X = W.mixture(n, dim, lists // 2, 5000 + lists)
    ix = ndb.IvfIndex(dim, lists)
    Cn = np.ascontiguousarray(X[:lists] + 0)           # centroids = data rows: exact zero distances, duplicates of clusters
    Cn[7] = Cn[3]                                      # two identical centroids: the lower index must win
    ix.set_centroids(Cn)
    got = ix.assign(X)
    want = orc.ivf_assign(X, Cn)
    assert np.array_equal(got, want), "%d of %d rows differ" % ((got != want).sum(), n)
    assert not np.any(got == 7)


def test_kmeans_on_the_tensor_cores_equals_the_oracle(ndb, orc):
    """kmeans_run (ivf_am.c:2117-2159) with k >= 256: assignment by squared L2 on the tensor cores, certified -- centroids,
    counts, iterations and cost stay bit-identical to the oracle (any wrong assignment would change the f32 sums)."""
    X = W.mixture(6000, 64, 300, 77)
    C, assign, counts, iters, cost = ndb.kmeans_train(X, 300)
    oC, oassign, ocounts, oiters, ocost = orc.kmeans_train(X, 300)
    assert iters == oiters and np.array_equal(assign, oassign) and np.array_equal(counts, ocounts)
    assert np.array_equal(BITS(C), BITS(oC)) and np.float32(cost) == np.float32(ocost)
