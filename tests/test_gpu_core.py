"""GPU parity: operator distances, exact kNN, k-means, through the C ABI, against the oracle."""
import os

import numpy as np
import pytest

import workloads as W

pytestmark = pytest.mark.gpu

BITS = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_bits(a, b):
    return np.array_equal(BITS(a), BITS(b))


@pytest.mark.parametrize("dim", [1, 2, 3, 4, 5, 7, 31, 32, 33, 96, 128, 131, 768])
@pytest.mark.parametrize("metric", [1, 2, 3])
def test_pairs_bit_exact_all_arith(ndb, orc, dim, metric):
    rng = np.random.default_rng(dim * 10 + metric)
    n = 777
    A = rng.standard_normal((n, dim)).astype(np.float32)
    B = rng.standard_normal((n, dim)).astype(np.float32)
    A[5] = 0.0                      # zero-norm row: cosine special cases
    B[6] = A[6]                     # identical vectors
    for arith, oarith in ((ndb.ARITH_OP_F64, orc.ARITH_OP_F64), (ndb.ARITH_IVF_F32, orc.ARITH_IVF_F32),
                          (ndb.ARITH_HNSW, orc.ARITH_HNSW)):
        got = ndb.distance_pairs(A, B, metric, arith)
        want = orc.distance_pairs(A, B, metric, oarith)
        assert same_bits(got, want), (dim, metric, arith, np.flatnonzero(BITS(got) != BITS(want))[:5])


def test_pairs_fast_within_1e5(ndb, orc):
    rng = np.random.default_rng(3)
    A = rng.standard_normal((4096, 128)).astype(np.float32)
    B = rng.standard_normal((4096, 128)).astype(np.float32)
    for metric in (1, 2):
        got = ndb.distance_pairs(A, B, metric, ndb.ARITH_FAST)
        want = orc.distance_pairs(A, B, metric, orc.ARITH_OP_F64)
        # north_star: fp32 paths within 1e-5 relative
        assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-30)) < 1e-5


def test_known_answers_t005(ndb):
    """NeuronDB/t/005_distances_comprehensive.t:40-56,140-158,248-255."""
    assert ndb.vector_l2_distance_op([0, 0], [3, 4])[0] == 5.0
    assert ndb.vector_l2_distance_op([1, 2, 3], [1, 2, 3])[0] == 0.0
    assert ndb.vector_l2_distance_op([0, 0, 0], [0, 0, 0])[0] == 0.0
    assert ndb.vector_cosine_distance_op([1, 0], [0, 1])[0] == 1.0
    assert ndb.vector_cosine_distance_op([1, 0], [1, 0])[0] == 0.0
    assert ndb.vector_cosine_distance_op([1, 0], [-1, 0])[0] == 2.0
    assert ndb.vector_inner_product_distance_op([1, 2, 3], [4, 5, 6])[0] == 32.0
    assert ndb.vector_inner_product_distance_op([1, 0], [0, 1])[0] == 0.0


def test_operator_errors(ndb):
    with pytest.raises(ndb.NdbError) as e:
        ndb.vector_l2_distance_op([1, 2], [1, 2, 3])
    assert e.value.code == -5
    with pytest.raises(ndb.NdbError) as e:
        ndb.vector_l2_distance_op([np.nan, 2], [1, 2])
    assert e.value.code == -4
    with pytest.raises(ndb.NdbError) as e:
        ndb.vector_l2_distance_op([np.inf, 2], [1, 2])
    assert e.value.code == -4


def test_backend_vtable_launchers(ndb, orc):
    rng = np.random.default_rng(11)
    A = rng.standard_normal((300, 64)).astype(np.float32)
    B = rng.standard_normal((300, 64)).astype(np.float32)
    got = ndb.launch_l2_distance(A, B)
    assert same_bits(got, orc.distance_pairs(A, B, 1, orc.ARITH_IVF_F32))
    cos = ndb.launch_cosine(A, B)
    want = orc.distance_pairs(A, B, 2, orc.ARITH_OP_F64)
    assert np.max(np.abs(cos - want)) < 1e-5


@pytest.mark.parametrize("n,dim,nq,k", [(1000, 128, 37, 10), (5000, 33, 64, 1), (20000, 128, 100, 10),
                                        (3000, 7, 9, 32), (4097, 96, 17, 100), (50, 16, 5, 10)])
@pytest.mark.parametrize("metric", [1, 2, 3])
def test_knn_exact_operator_arith(ndb, orc, n, dim, nq, k, metric):
    X = W.gaussian(n, dim, 1234 + n)
    Q = W.gaussian(nq, dim, 4321 + n)
    ds = ndb.Dataset(dim)
    ds.append(X[: n // 2])
    ds.append(X[n // 2:])           # appended in two pieces: exercises partial last blocks
    assert len(ds) == n
    d, i = ds.knn(Q, k, metric, ndb.ARITH_OP_F64)
    od, oi = orc.knn_exact(X, Q, k, metric, orc.ARITH_OP_F64)
    assert np.array_equal(i, oi)
    assert same_bits(d, od)
    ds.close()


def test_knn_exact_ties_and_duplicates(ndb, orc):
    rng = np.random.default_rng(5)
    base = rng.integers(-2, 3, size=(64, 8)).astype(np.float32)
    X = np.concatenate([base] * 40)         # every vector 40 times: massive exact ties
    Q = base[:16].copy()
    ds = ndb.Dataset(8)
    ds.append(X)
    d, i = ds.knn(Q, 10, 1, ndb.ARITH_OP_F64)
    od, oi = orc.knn_exact(X, Q, 10, 1, orc.ARITH_OP_F64)
    assert np.array_equal(i, oi) and same_bits(d, od)


def test_knn_exact_fewer_rows_than_k(ndb, orc):
    X = W.gaussian(6, 12, 1)
    Q = W.gaussian(3, 12, 2)
    ds = ndb.Dataset(12)
    ds.append(X, ids=np.arange(100, 106))
    d, i = ds.knn(Q, 10, 1, ndb.ARITH_OP_F64)
    od, oi = orc.knn_exact(X, Q, 10, 1, orc.ARITH_OP_F64, ids=np.arange(100, 106))
    assert np.array_equal(i, oi)
    assert np.all(i[:, 6:] == -1) and np.all(np.isinf(d[:, 6:]))
    assert same_bits(d[:, :6], od[:, :6])


def test_knn_exact_config1_slice(ndb, orc):
    """BASELINE config 1 shape (100k x 128, k=10) on a 64-query slice against the oracle."""
    X = W.gaussian(100_000, 128, 1234)
    Q = W.gaussian(1000, 128, 4321)
    ds = ndb.Dataset(128)
    ds.append(X)
    d, i = ds.knn(Q, 10, 1, ndb.ARITH_OP_F64)
    od, oi = orc.knn_exact(X, Q[:64], 10, 1, orc.ARITH_OP_F64)
    assert np.array_equal(i[:64], oi) and same_bits(d[:64], od)
    # size-independent properties at full size: sorted, ids valid and unique per query
    assert np.all(np.diff(d, axis=1) >= 0)
    assert np.all((i >= 0) & (i < 100_000))
    assert all(len(set(r)) == 10 for r in i)
    # FAST path: same ids wherever the gap to the next distance exceeds the tolerance
    df, i_f = ds.knn(Q, 10, 1, ndb.ARITH_FAST)
    assert np.max(np.abs(df - d) / d) < 1e-5
    assert (i_f == i).mean() > 0.999


@pytest.mark.parametrize("n,dim,k", [(2000, 16, 7), (10000, 128, 64), (513, 5, 512), (3000, 33, 100)])
def test_kmeans_train_literal(ndb, orc, n, dim, k):
    X = W.mixture(n, dim, max(2, k // 2), 77 + n)
    C, assign, counts, iters, cost = ndb.kmeans_train(X, k)
    oC, oassign, ocounts, oiters, ocost = orc.kmeans_train(X, k)
    assert iters == oiters
    assert np.array_equal(assign, oassign)
    assert np.array_equal(counts, ocounts)
    assert same_bits(C, oC)
    assert np.float32(cost).view(np.uint32) == np.float32(ocost).view(np.uint32)


def test_kmeans_vtable(ndb, orc):
    X = W.gaussian(5000, 24, 9)
    C0 = X[:50].copy()
    idx = ndb.launch_kmeans_assign(X, C0)
    assert np.array_equal(idx, orc.kmeans_assign(X, C0))
    C1 = ndb.launch_kmeans_update(X, idx, 50)
    oC1, _ = orc.kmeans_update(X, idx, 50)
    assert same_bits(C1, oC1)


def test_merge_topk(ndb, orc):
    rng = np.random.default_rng(8)
    d = np.sort(rng.integers(0, 20, size=(4, 50, 10)).astype(np.float32), axis=2)
    ids = rng.permutation(4 * 50 * 10).reshape(4, 50, 10).astype(np.int64)
    ids[3, :, 7:] = -1
    d[3, :, 7:] = np.inf
    gd, gi = ndb.merge_topk(d, ids)
    od, oi = orc.merge_topk(d, ids)
    assert np.array_equal(gi, oi) and same_bits(gd, od)


# ---- key extraction ------------------------------------------------------
def test_keys_from_halfvec_every_half_bit_exact(ndb, orc):
    h = np.arange(65536, dtype=np.uint16).reshape(64, 1024)
    got = ndb.keys_from_halfvec(h)
    want = orc.keys_from_halfvec(h)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "fp16_table.npz"))["bits"]
    assert np.array_equal(got.view(np.uint32).reshape(-1), golden)          # the reference's own table


@pytest.mark.parametrize("n,nbits", [(1, 1), (33, 19), (500, 128), (3, 32767)])
def test_keys_from_bits(ndb, orc, n, nbits):
    bits = np.random.default_rng(nbits).integers(0, 256, size=(n, (nbits + 7) // 8), dtype=np.uint8)
    assert np.array_equal(ndb.keys_from_bits(bits, nbits), orc.keys_from_bits(bits, nbits))


def test_keys_from_sparse(ndb, orc):
    rng = np.random.default_rng(77)
    n, dim = 300, 257
    nnz = rng.integers(0, 40, size=n)
    nnz[5] = 0
    indptr = np.concatenate([[0], np.cumsum(nnz)]).astype(np.int64)
    indices = rng.integers(-3, dim + 3, size=int(indptr[-1])).astype(np.int32)     # some out of range, some repeated
    values = rng.standard_normal(int(indptr[-1])).astype(np.float32)
    got = ndb.keys_from_sparse(indptr, indices, values, dim)
    assert np.array_equal(got, orc.keys_from_sparse(indptr, indices, values, dim))
    # the extracted rows feed the index like any other
    ix = ndb.IvfIndex(dim, 4)
    ix.ivfbuild(got)
    ix.ivfinsert(got)
    d, i = ix.search(got[:5], 4, 1)
    assert np.array_equal(i[:, 0], np.arange(5)) and np.all(d[:, 0] == 0)


def test_keys_reject_bad_dimensions(ndb):
    with pytest.raises(ndb.NdbError):
        ndb.keys_from_bits(np.zeros((1, 4096), np.uint8), 32768)
    with pytest.raises(ndb.NdbError):
        ndb.keys_from_sparse(np.array([0, 0], np.int64), np.zeros(0, np.int32), np.zeros(0, np.float32), 0)


def test_keys_from_vector_datums(ndb):
    """struct Vector (neurondb.h:35-41): int32 vl_len_ (4-byte varlena header: size << 2), int16 dim, int16 unused, data."""
    rng = np.random.default_rng(8)
    n, dim = 37, 20
    X = rng.standard_normal((n, dim)).astype(np.float32)
    hdr = np.zeros((n, 2), np.int32)
    hdr[:, 0] = (8 + 4 * dim) << 2
    hdr[:, 1] = dim                                   # int16 dim | int16 unused (little endian)
    datums = np.concatenate([hdr.view(np.uint8).reshape(n, 8), X.view(np.uint8).reshape(n, 4 * dim)], axis=1)
    assert np.array_equal(ndb.keys_from_vector(datums, dim), X)
    datums[5, 4] = dim + 1                            # a datum of another dimension
    with pytest.raises(ndb.NdbError) as e:
        ndb.keys_from_vector(datums, dim)
    assert "datum 5" in str(e.value)


@pytest.mark.parametrize("metric", [1, 2, 3])
def test_operator_distances_as_an_avx_build_computes_them(ndb, orc, metric):
    """8 / 16 f32 lane accumulators, fixed reduction tree, scalar tail -- bit for bit
    against the oracle's lane-by-lane restatement and, for AVX2, the reference's own sources built
    with -mavx2 -mfma under the shim."""
    rng = np.random.default_rng(metric)
    for dim in list(range(1, 41)) + [63, 64, 65, 127, 128, 129, 768, 1000]:
        A = rng.standard_normal((67, dim)).astype(np.float32)
        B = rng.standard_normal((67, dim)).astype(np.float32)
        if dim > 3:
            A[5] = 0.0                                        # zero norm: cosine returns 1.0
        for arith, oarith in ((ndb.ARITH_AVX2, orc.ARITH_AVX2), (ndb.ARITH_AVX512, orc.ARITH_AVX512)):
            got = ndb.distance_pairs(A, B, metric, arith)
            want = orc.distance_pairs(A, B, metric, oarith)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (dim, arith)
        if orc.ref_lib(avx2=True) is not None:
            ref = orc.ref_distance_pairs(metric, A, B, avx2=True)
            assert np.array_equal(ndb.distance_pairs(A, B, metric, ndb.ARITH_AVX2).view(np.uint32), ref.view(np.uint32)), dim
    # one query against rows (vector_*_distance_batch)
    X = rng.standard_normal((100, 24)).astype(np.float32)
    q = rng.standard_normal(24).astype(np.float32)
    got = ndb.distance_rows(X, q, metric, ndb.ARITH_AVX2)
    want = orc.distance_pairs(X, np.repeat(q[None], 100, 0), metric, orc.ARITH_AVX2)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_index_arithmetic_equals_the_reference_golden_vectors(ndb):
    """tests/golden/index_leafs.npz holds outputs of the reference's OWN ivfComputeDistance,
    hnswComputeDistance and k-means block (cut out of ivf_am.c / hnsw_am.c and compiled,
    oracle/extract_ref_leafs.py): the kernels must reproduce them bit for bit, no oracle in between."""
    import test_oracle as T
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "index_leafs.npz"))
    ivf, hn = [], []
    for a, b in T._leaf_inputs():
        for s in (1, 2):
            ivf.append(ndb.distance_pairs(a, b, s, ndb.ARITH_IVF_F32))
        for s in (1, 2, 3):
            hn.append(ndb.distance_pairs(a, b, s, ndb.ARITH_HNSW))
    assert np.array_equal(np.concatenate(ivf).view(np.uint32), g["ivf_bits"])
    assert np.array_equal(np.concatenate(hn).view(np.uint32), g["hnsw_bits"])
    X = W.mixture(1500, 16, 10, 31)
    C, assign, counts, iters, cost = ndb.kmeans_train(X, 24)
    assert np.array_equal(C.view(np.uint32), g["km_C_bits"])
    assert np.array_equal(assign, g["km_assign"]) and np.array_equal(counts, g["km_counts"])


def test_reference_index_fixture_t010(ndb):
    """t/010_indexes_comprehensive.t:32-48 -- 8 rows [1+i, 2+i, 3+i, 4+i], query [1,2,3,4]: row i at distance 2 i,
    through the exact scan, an IVF index (vector_l2_ops) and an HNSW index with the default options."""
    import test_oracle as T
    X, q = T.FIXTURE_T010, T.FIXTURE_T010[:1]
    want_d = np.arange(8, dtype=np.float32) * 2
    ds = ndb.Dataset(4)
    ds.append(X)
    d, i = ds.knn(q, 8, ndb.L2, ndb.ARITH_OP_F64)
    assert np.array_equal(i[0], np.arange(8)) and np.array_equal(d[0], want_d)
    ix = ndb.IvfIndex(4, 4)
    ix.ivfbuild(X)
    ix.ivfinsert(X)
    d, i = ix.search(q, 4, 8)
    assert np.array_equal(i[0], np.arange(8)) and np.array_equal(d[0], want_d)
    h = ndb.HnswIndex(4, 16, 200, 64)
    h.hnswbuild(X, levels=np.zeros(8, np.int32))
    d, i = h.search(q, 64, 8)
    assert np.array_equal(i[0], np.arange(8)) and np.array_equal(d[0], want_d)
