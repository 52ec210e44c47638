"""GPU: the SQL-function semantics routed through the ABI: vector_*_distance_batch
(src/vector/vector_batch.c:37-420), ivf_knn_search_gpu / hnsw_knn_search_gpu (src/gpu/common/gpu_sql.c:498-1456)."""
import numpy as np
import pytest

import workloads as W

pytestmark = pytest.mark.gpu
BITS = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("metric", [1, 2, 3])
def test_vector_distance_batch_nulls_and_operator_arithmetic(ndb, orc, metric):
    rng = np.random.default_rng(metric)
    q = rng.standard_normal(48).astype(np.float32)
    vecs = [rng.standard_normal(48).astype(np.float32) for _ in range(200)]
    vecs[3] = None                                       # NULL element -> NULL
    vecs[10] = rng.standard_normal(47).astype(np.float32)  # other dimension -> NULL (vector_batch.c:144-150)
    vecs[11] = np.zeros(48, np.float32)                  # zero vector: cosine distance 1.0
    d, nulls = ndb.vector_distance_batch(vecs, q, metric)
    assert nulls[3] and nulls[10] and nulls.sum() == 2
    ok = [i for i in range(200) if not nulls[i]]
    A = np.stack([vecs[i] for i in ok])
    want = orc.distance_pairs(A, np.repeat(q[None], len(ok), 0), metric, orc.ARITH_OP_F64)     # the operators' fp64 loops
    assert np.array_equal(BITS(d[ok]), BITS(want))
    if metric == 3:                                      # -inner_product_distance = +dot (:404)
        assert abs(d[ok[0]] - float(A[0].astype(np.float64) @ q)) < 1e-4
    # all elements valid: the one-launch path
    d2, n2 = ndb.vector_distance_batch([vecs[i] for i in ok], q, metric)
    assert not n2.any() and np.array_equal(BITS(d2), BITS(want))
    with pytest.raises(ndb.NdbError):                    # "vector array must not be empty"
        ndb.vector_distance_batch([], q, metric)


def test_knn_search_gpu_functions(ndb, orc):
    X = W.mixture(6000, 24, 20, 31)
    Q = W.mixture(10, 24, 20, 32, centers_seed=31)
    ix = ndb.IvfIndex(24, 20)
    ix.ivfbuild(X)
    lists = ix.ivfinsert(X)
    off, rows = orc.lists_from_assignment(lists, 20)
    od, oi, _ = orc.ivf_search(X, ix.centroids(), off, rows, Q, 5, 7)
    for q in range(10):
        ids, d = ix.knn_search_gpu(Q[q], 7, 5)
        assert np.array_equal(ids, oi[q]) and np.array_equal(BITS(d), BITS(od[q]))
    for bad in (dict(k=0), dict(k=10001), dict(nprobe=0), dict(nprobe=1001)):       # gpu_sql.c:983-991
        with pytest.raises(ndb.NdbError) as e:
            ix.knn_search_gpu(Q[0], bad.get("k", 5), bad.get("nprobe", 5))
        assert e.value.code == -1
    with pytest.raises(ndb.NdbError) as e:
        ix.knn_search_gpu(Q[0][:20], 5, 5)
    assert e.value.code == -5
    levels = orc.hnsw_levels(1500, seed=2)
    h = ndb.HnswIndex(24, 8, 32, 32)
    h.hnswbuild(X[:1500], levels=levels, batch=1)
    og = orc.Hnsw(24, 8, 32, 32, capacity=1500)
    og.build(X[:1500], levels, 1)
    od, on, _ = og.search(Q, 64, 5, 1, 1)
    for q in range(10):
        ids, d = h.knn_search_gpu(Q[q], 5, 64)
        assert np.array_equal(ids, on[q].astype(np.int64)) and np.array_equal(BITS(d), BITS(od[q]))
    with pytest.raises(ndb.NdbError):
        h.knn_search_gpu(Q[0], 5, 0)                     # ef_search must be between 1 and 10000
