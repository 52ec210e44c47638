"""GPU: product quantisation (NeuronDB/src/ml/ml_product_quantization.c) through the C ABI -- codebook training
(train_pq_codebook / train_subspace_kmeans), encoding (pq_encode_vector, the vtable's launch_pq_encode) and the
asymmetric-distance scan (pq_asymmetric_distance) with its top-k.

Codebooks (bit patterns), codes and every float distance must equal (a) the committed outputs of the reference's OWN
code (tests/golden/ml_paths.npz: train_subspace_kmeans compiled from its source, the loops of pq_encode_vector /
pq_asymmetric_distance cut out as text; tests/golden/make_golden.py) and (b) the oracle restatement on larger inputs."""
import os

import numpy as np
import pytest

import oracle_lib as O
import workloads as W

pytestmark = pytest.mark.gpu
BITS = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ml_paths.npz")
CASES = [(600, 16, 4, 16, 11), (900, 24, 8, 32, 12), (300, 12, 3, 256, 13), (50, 8, 8, 2, 14)]       # test_oracle._pq_cases


def test_pq_equals_the_reference_outputs(ndb):
    g = np.load(GOLDEN)
    for n, dim, m, ksub, seed in CASES:
        tag = "pq%d" % n
        X = W.mixture(n, dim, 6, seed)
        Q = W.mixture(12, dim, 6, seed + 1, centers_seed=seed)
        cb = ndb.pq_train(X, m, ksub, g["draws_" + tag], 5)
        assert np.array_equal(BITS(cb), g["cb_bits_" + tag]), tag
        codes = ndb.pq_encode(X, cb)
        assert np.array_equal(codes, g["codes_" + tag]), tag
        assert np.array_equal(ndb.launch_pq_encode(X, cb).astype(np.int16), codes)
        pq = ndb.PqIndex(cb)
        assert np.array_equal(pq.add(X, want_codes=True), codes) and len(pq) == n
        d, _ = pq.distances(Q)
        assert np.array_equal(BITS(d), g["adc_bits_" + tag]), tag


@pytest.mark.parametrize("n,dim,m,ksub,k", [(70000, 32, 8, 256, 10), (5000, 30, 5, 64, 100), (9000, 128, 16, 256, 32), (40, 6, 6, 3, 50)])
def test_pq_scan_equals_the_oracle(ndb, n, dim, m, ksub, k):
    X = W.mixture(n, dim, 12, n + m)
    Q = W.mixture(24, dim, 12, n + m + 1, centers_seed=n + m)
    draws = np.random.default_rng(n).integers(0, O.RAND_MAX, m * ksub, dtype=np.int64).astype(np.int32)
    ntrain = min(n, 2000)
    cb = ndb.pq_train(X[:ntrain], m, ksub, draws, 3)
    assert np.array_equal(BITS(cb), BITS(O.pq_train(X[:ntrain], m, ksub, draws, 3)))
    pq = ndb.PqIndex(cb)
    half = n // 2
    codes = np.concatenate([pq.add(X[:half], want_codes=True), pq.add(X[half:], want_codes=True)])
    want_codes = O.pq_encode(X, cb)
    assert np.array_equal(codes, want_codes)
    wd, wr, wall = O.pq_knn(Q, want_codes, cb, k, want_all=True)
    d, r = pq.search(Q, k)
    assert np.array_equal(BITS(d), BITS(wd)) and np.array_equal(r, wr)           # +inf / -1 past the end included
    da, rechecked = pq.distances(Q)
    assert np.array_equal(BITS(da), BITS(wall))
    assert rechecked <= max(4, da.size // 1000)                                  # the chain re-evaluation is the exception
    # codes made elsewhere (the SQL function's int2[]) give the same scan
    pq2 = ndb.PqIndex(cb)
    pq2.add_codes(want_codes)
    d2, r2 = pq2.search(Q, k)
    assert np.array_equal(BITS(d2), BITS(wd)) and np.array_equal(r2, wr)


def test_pq_scan_recheck_path_is_exact(ndb):
    """The certificate sends a distance to the reference's chain when the table sum lies within the error band of a float
    rounding boundary.  NDB_PQ_EPS widens the band (1e-9 relative against a float spacing of about 1e-7: a few per cent of the
    distances fall in it; with 1e-1 all of them do), the returned floats must not change."""
    X = W.mixture(6000, 24, 8, 77)
    Q = W.mixture(16, 24, 8, 78, centers_seed=77)
    draws = np.random.default_rng(7).integers(0, O.RAND_MAX, 6 * 64, dtype=np.int64).astype(np.int32)
    cb = ndb.pq_train(X[:1500], 6, 64, draws, 2)
    pq = ndb.PqIndex(cb)
    codes = pq.add(X, want_codes=True)
    wd, wr, wall = O.pq_knn(Q, codes, cb, 10, want_all=True)
    seen = []
    try:
        for eps in ("1e-9", "1e-1"):
            os.environ["NDB_PQ_EPS"] = eps
            da, rechecked = pq.distances(Q)
            assert np.array_equal(BITS(da), BITS(wall))
            d, r = pq.search(Q, 10)
            assert np.array_equal(BITS(d), BITS(wd)) and np.array_equal(r, wr)
            seen.append(rechecked)
    finally:
        del os.environ["NDB_PQ_EPS"]
    assert 0 < seen[0] < seen[1] and seen[1] > da.size // 2


def test_pq_scan_float_screen_changes_nothing(ndb):
    """The top-k scan screens rows with a float copy of the table against the block's best k-th distance; NDB_PQ_NO_FILTER
    sends every row down the exact path.  Same distances and rows either way -- also when the rows arrive in descending
    distance (every row passes the screen), with duplicates (ties at the k-th place) and with tiny distances."""
    m, ksub, dim = 8, 256, 32
    X = W.mixture(40000, dim, 10, 5)
    draws = np.random.default_rng(8).integers(0, O.RAND_MAX, m * ksub, dtype=np.int64).astype(np.int32)
    cb = ndb.pq_train(X[:3000], m, ksub, draws, 3)
    codes = O.pq_encode(X, cb)
    Q = W.mixture(12, dim, 10, 6, centers_seed=5)
    Q[1] = X[7]
    _, _, al = O.pq_knn(Q[:1], codes, cb, 1, want_all=True)
    order = np.argsort(-al[0], kind="stable")                        # descending distance to query 0
    codes_desc = np.ascontiguousarray(codes[order])
    codes_desc[100:160] = codes_desc[39990]                          # 60 copies of a near row: ties at the k-th place
    cb_tiny = (cb * np.float32(1e-20)).astype(np.float32)            # squared distances around 1e-40: float subnormals
    for cbk, cd, qs in ((cb, codes_desc, Q), (cb_tiny, codes_desc, (Q * np.float32(1e-20)).astype(np.float32))):
        for k in (10, 70):
            wd, wr, _ = O.pq_knn(qs, cd, cbk, k)
            got = {}
            for mode in ("screen", "exact"):
                if mode == "exact":
                    os.environ["NDB_PQ_NO_FILTER"] = "1"
                try:
                    pq = ndb.PqIndex(cbk)
                    pq.add_codes(cd)
                    got[mode] = pq.search(qs, k)
                finally:
                    os.environ.pop("NDB_PQ_NO_FILTER", None)
                assert np.array_equal(BITS(got[mode][0]), BITS(wd)) and np.array_equal(got[mode][1], wr), (mode, k)


def test_pq_errors_are_the_sql_functions(ndb):
    X = W.gaussian(64, 12, 1)
    draws = np.arange(4096, dtype=np.int32)
    with pytest.raises(ndb.NdbError) as e:
        ndb.pq_train(X, 0, 16, draws[:0])
    assert e.value.code == -1 and "m (number of subspaces) must be" in str(e.value)
    with pytest.raises(ndb.NdbError) as e:
        ndb.pq_train(X, 4, 1, draws[:4])
    assert e.value.code == -1 and "ksub (centroids per subspace) must be" in str(e.value)
    with pytest.raises(ndb.NdbError) as e:
        ndb.pq_train(X, 5, 4, draws[:20])
    assert e.value.code == -5 and "Vector dimension 12 must be divisible by number of subspaces m=5" in str(e.value)
    cb = ndb.pq_train(X, 4, 8, draws[:32], 2)
    pq = ndb.PqIndex(cb)
    with pytest.raises(ndb.NdbError) as e:
        pq.search(X[:2], 3)                                   # nothing encoded yet
    assert e.value.code == -7
    bad = np.zeros((3, 4), np.int16)
    bad[1, 2] = 8
    with pytest.raises(ndb.NdbError) as e:
        pq.add_codes(bad)
    assert e.value.code == -8 and "Invalid PQ code 8 at subspace 2 (valid: 0-7)" in str(e.value) and len(pq) == 0
    pq.add(X)
    q = X[:2].copy()
    q[0, 0] = np.nan
    with pytest.raises(ndb.NdbError) as e:
        pq.search(q, 3)
    assert e.value.code == -4
