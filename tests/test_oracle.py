"""CPU: the oracle against the reference's known answers, the reference's own compiled operator
sources (when present) and the committed golden vectors."""
import importlib.util
import os

import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
import workloads as W

HERE = os.path.dirname(os.path.abspath(__file__))
BITS = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)


def _golden_mod():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def d1(fn, a, b, *extra):
    a, b = O.f32(a), O.f32(b)
    return fn(a, b, a.size, *extra)


def test_known_answers_t005():
    """NeuronDB/t/005_distances_comprehensive.t:40-56,140-158,248-255."""
    l = O.lib()
    assert d1(l.orc_l2_distance, [0, 0], [3, 4]) == 5.0
    assert d1(l.orc_l2_distance, [1, 2, 3], [1, 2, 3]) == 0.0
    assert d1(l.orc_l2_distance, [0, 0], [0, 0]) == 0.0
    assert d1(l.orc_cosine_distance, [1, 0], [0, 1]) == 1.0
    assert d1(l.orc_cosine_distance, [1, 0], [1, 0]) == 0.0
    assert d1(l.orc_cosine_distance, [1, 0], [-1, 0]) == 2.0
    assert d1(l.orc_inner_product_op, [1, 2, 3], [4, 5, 6]) == 32.0          # <#> yields +dot
    assert d1(l.orc_inner_product_distance, [1, 2, 3], [4, 5, 6]) == -32.0
    assert d1(l.orc_inner_product_op, [1, 0], [0, 1]) == 0.0
    assert d1(l.orc_cosine_distance, [0, 0], [1, 0]) == 1.0                  # vector_distance.c:201-202
    assert d1(l.orc_hnsw_distance, [0, 0], [1, 0], 2) == 2.0                 # hnsw_am.c:1329-1330 (Q15)
    assert d1(l.orc_ivf_distance, [0, 0], [1, 0], 2) == 1.0
    assert l.orc_check_vector(O.f32([1, np.nan, 2]), 3) == 1
    assert l.orc_check_vector(O.f32([1, 2, np.inf]), 3) == 2
    assert l.orc_check_vector(O.f32([1, 2, 3]), 3) == -1


def test_golden_operator_distances():
    """Outputs of the reference's own fmgr functions, committed as fixtures."""
    g = np.load(os.path.join(HERE, "golden", "operator_distances.npz"))
    mg = _golden_mod()
    for dim in mg.DIMS:
        A, B = mg.operator_inputs(dim)
        for metric in (1, 2, 3):
            got = O.distance_pairs(A, B, metric, O.ARITH_OP_F64)
            assert np.array_equal(BITS(got), BITS(g["scalar_m%d_d%d" % (metric, dim)])), (dim, metric)
            got = O.distance_pairs(A, B, metric, O.ARITH_AVX2)
            assert np.array_equal(BITS(got), BITS(g["avx2_m%d_d%d" % (metric, dim)])), (dim, metric, "avx2")


@pytest.mark.skipif(O.ref_lib() is None, reason="oracle/_ref not built (reference tree absent)")
def test_oracle_equals_reference_build_bit_for_bit():
    rng = np.random.default_rng(42)
    for dim in [1, 5, 8, 13, 16, 29, 64, 128, 200, 768, 1536]:
        A = (rng.standard_normal((300, dim)) * rng.choice([1e-3, 1.0, 1e3])).astype(np.float32)
        B = rng.standard_normal((300, dim)).astype(np.float32)
        for metric in (1, 2, 3):
            assert np.array_equal(BITS(O.ref_distance_pairs(metric, A, B)),
                                  BITS(O.distance_pairs(A, B, metric, O.ARITH_OP_F64))), (dim, metric)
            assert np.array_equal(BITS(O.ref_distance_pairs(metric, A, B, avx2=True)),
                                  BITS(O.distance_pairs(A, B, metric, O.ARITH_AVX2))), (dim, metric, "avx2")


@pytest.mark.skipif(O.ref_lib() is None, reason="oracle/_ref not built (reference tree absent)")
def test_reference_error_behaviour():
    """dimension mismatch / NaN / Inf are rejected (t/005:106-132,285-300; vector_distance.c:34-74)."""
    rc, _, msg = O.ref_distance(1, [1, 2], [1, 2, 3])
    assert rc != 0 and "dimension" in msg
    rc, _, msg = O.ref_distance(1, [np.nan, 1], [1, 2])
    assert rc != 0 and "NaN" in msg
    rc, _, msg = O.ref_distance(2, [np.inf, 1], [1, 2])
    assert rc != 0
    rc, v, _ = O.ref_distance(1, [0, 0], [3, 4])
    assert rc == 0 and v == 5.0


def test_golden_index_paths():
    g = np.load(os.path.join(HERE, "golden", "index_paths.npz"))
    X = W.mixture(3000, 16, 12, 5)
    Q = W.mixture(40, 16, 12, 6, centers_seed=5)
    C, assign, counts, iters, cost = O.kmeans_train(X[:1200], 12)
    assert np.array_equal(BITS(C), BITS(g["km_C"])) and np.array_equal(assign, g["km_assign"])
    assert iters == int(g["km_iters"]) and np.float32(cost) == g["km_cost"]
    lists = O.ivf_assign(X, C)
    assert np.array_equal(lists, g["ivf_lists"])
    off, rows = O.lists_from_assignment(lists, 12)
    for lit in (0, 1):
        for metric in (1, 2, 3):
            d, i, _ = O.ivf_search(X, C, off, rows, Q, 4, 10, strategy=metric, literal=bool(lit))
            assert np.array_equal(i, g["ivf_i_l%d_m%d" % (lit, metric)])
            assert np.array_equal(BITS(d), BITS(g["ivf_d_l%d_m%d" % (lit, metric)]))
    levels = O.hnsw_levels(1500, seed=9)
    assert np.array_equal(levels, g["hnsw_levels"])
    for mode in (0, 1, 3):
        h = O.Hnsw(16, 6, 24, 24, capacity=1500)
        h.build(X[:1500], levels, mode)
        e = h.export()
        assert np.array_equal(e["nbr0"], g["hnsw_nbr0_b%d" % mode])
        for smode in (0, 1):
            d, n, _ = h.search(Q, 24, 10, 1, smode)
            assert np.array_equal(n, g["hnsw_n_b%d_s%d" % (mode, smode)])
    d, i = O.knn_exact(X, Q, 10, 1, O.ARITH_OP_F64)
    assert np.array_equal(i, g["knn_i"]) and np.array_equal(BITS(d), BITS(g["knn_d"]))


def test_kmeans_literal_semantics():
    """centroids := first k samples (Q8); empty clusters stay at zero; strict < keeps the lowest index."""
    X = np.zeros((40, 3), np.float32)
    X[:, 0] = np.arange(40) % 4                     # four distinct points, many duplicates
    C, assign, counts, iters, cost = O.kmeans_train(X, 6, max_iter=50, threshold=0.001)
    # samples 4,5 duplicate samples 0,1: centroids 4,5 start equal to 0,1 and lose every tie
    assert counts[4] == 0 and counts[5] == 0
    assert np.all(C[4] == 0) and np.all(C[5] == 0)
    assert sorted(np.unique(assign).tolist()) == [0, 1, 2, 3]
    assert cost == 0.0 and iters == 2               # second pass: |prev - cost| = 0 < 0.001
    assert O.lib().orc_ivf_train_samples(1_000_000, 1024) == 10000
    assert O.lib().orc_ivf_train_samples(1_000_000, 50) == 5000
    assert O.lib().orc_ivf_train_samples(300, 50) == 300


def test_ivf_literal_cap_and_ties():
    """k*10 candidate cap in probe order (Q9); select_clusters ties go to the lowest list id."""
    X = W.gaussian(2000, 8, 3)
    C = X[:4].copy()
    C[1] = C[0]                                      # two identical centroids
    probes = O.select_clusters(X[0], C, 4)
    assert probes[0] == 0 and probes[1] == 1         # equal distance: lower index first
    lists = O.ivf_assign(X, C)
    assert not np.any(lists == 1)                    # strict <: the duplicate never wins
    off, rows = O.lists_from_assignment(lists, 4)
    d_lit, i_lit, c_lit = O.ivf_search(X, C, off, rows, X[:20], 4, 10, literal=True)
    d_full, i_full, _ = O.ivf_search(X, C, off, rows, X[:20], 4, 10, literal=False)
    assert np.all(c_lit == 10)
    # the capped scan only ever sees the first 100 entries in probe order
    for q in range(20):
        first = []
        for l in O.select_clusters(X[q], C, 4):
            first += rows[off[l]:off[l + 1]].tolist()
        assert set(i_lit[q]) <= set(first[:100])
    assert np.all(d_full[:, 0] == 0.0) and np.all(i_full[:, 0] == np.arange(20))
    assert np.all(d_full <= d_lit + 1e-30)


def test_knn_ties_by_id_and_recall():
    X = np.tile(np.arange(8, dtype=np.float32)[:, None], (3, 4))        # rows 0..7 three times
    ids = np.arange(24)[::-1].copy()                                     # descending ids
    d, i = O.knn_exact(X, X[:2], 6, 1, O.ARITH_OP_F64, ids=ids)
    assert np.all(d[0, :3] == 0) and i[0, :3].tolist() == sorted(i[0, :3].tolist())
    assert O.recall_at_k(np.array([[1, 2, 3, 4]]), np.array([[4, 3, 9, 8]])) == 0.5


def test_merge_topk_matches_lexsort():
    rng = np.random.default_rng(0)
    d = rng.integers(0, 9, size=(3, 20, 5)).astype(np.float32)
    ids = rng.permutation(300).reshape(3, 20, 5).astype(np.int64)
    md, mi = O.merge_topk(d, ids)
    for q in range(20):
        dd = d[:, q, :].reshape(-1); ii = ids[:, q, :].reshape(-1)
        order = np.lexsort((ii, dd))[:5]
        assert np.array_equal(mi[q], ii[order]) and np.array_equal(md[q], dd[order])


def test_hnsw_oracle_properties():
    X = W.gaussian(1200, 12, 21)
    Q = W.gaussian(50, 12, 22)
    levels = O.hnsw_levels(1200, seed=4)
    assert levels.min() == 0 and levels.max() < 16 and 0.02 < (levels > 0).mean() < 0.12   # P(level>=1)=e^(-1/0.36)
    gt = W.exact_ground_truth(X, Q, 10)
    g = O.Hnsw(12, 8, 40, 40, capacity=1200)
    g.build(X, levels, 1)
    d, n, cnt = g.search(Q, 64, 10, 1, 1)
    assert np.all(cnt == 10) and np.all(np.diff(d, axis=1) >= 0)
    assert O.recall_at_k(n.astype(np.int64), gt) > 0.9
    # the literal BFS stops at ef candidates: never more than ef + 2m evaluations at level 0
    g.search(Q, 16, 10, 1, 0)
    lit_evals = g.distance_evals()
    g.search(Q, 16, 10, 1, 1)
    assert lit_evals < g.distance_evals()
    e = g.export()
    assert np.all(e["cnt"] <= 16) and e["nbr0"].shape == (1200, 16)
    assert np.all((e["nbr0"] == 0xFFFFFFFF) | (e["nbr0"] < 1200))


# ---- key extraction ------------------------------------------------------
@pytest.mark.skipif(O.ref_fp16_lib() is None, reason="oracle/_ref not built (reference tree absent)")
def test_fp16_restatement_equals_reference_for_every_half():
    ref = O.ref_fp16_lib()
    h = np.arange(65536, dtype=np.uint16)
    want = np.array([ref.fp16_to_float(int(x)) for x in h], np.float32)
    got = O.keys_from_halfvec(h.reshape(1, -1))[0]
    # (a float returned through ctypes passes through a Python double, which quiets signalling NaNs:
    # compare NaNs as NaNs, everything else bit for bit)
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan)
    assert np.array_equal(got[~nan].view(np.uint32), want[~nan].view(np.uint32))


def test_fp16_is_ieee_except_for_the_subnormal_quirk():
    """Pins the restatement without the reference tree: IEEE binary16 -> binary32 everywhere except
    subnormal halves, which the reference scales by a further 2^-10 (quantization.c:186-197)."""
    h = np.arange(65536, dtype=np.uint16)
    got = O.keys_from_halfvec(h.reshape(256, 256)).reshape(-1)
    ieee = h.view(np.float16).astype(np.float32)
    sub = ((h & 0x7C00) == 0) & ((h & 0x03FF) != 0)
    nan = np.isnan(ieee)
    assert np.array_equal(got[~sub & ~nan].view(np.uint32), ieee[~sub & ~nan].view(np.uint32))
    assert np.all(np.isnan(got[nan]))
    assert np.array_equal(got[sub], ieee[sub] * np.float32(2.0 ** -10))


def test_bit_and_sparse_keys():
    rng = np.random.default_rng(5)
    bits = rng.integers(0, 256, size=(7, 3), dtype=np.uint8)
    rows = O.keys_from_bits(bits, 19)                       # 19 bits: the last byte is partly used
    want = np.where(np.unpackbits(bits, axis=1)[:, :19] == 1, 1.0, -1.0).astype(np.float32)
    assert np.array_equal(rows, want)
    # sparsevec: out-of-range indices ignored, a repeated index keeps the last value, empty row = zeros
    indptr = np.array([0, 3, 3, 6], np.int64)
    indices = np.array([1, 4, 1, -1, 5, 2], np.int32)
    values = np.array([1.5, 2.5, 3.5, 9.0, 9.0, -4.0], np.float32)
    rows = O.keys_from_sparse(indptr, indices, values, 5)
    assert np.array_equal(rows, np.array([[0, 3.5, 0, 0, 2.5], [0, 0, 0, 0, 0], [0, 0, -4.0, 0, 0]], np.float32))


def test_fp16_restatement_equals_golden_table():
    """tests/golden/fp16_table.npz holds the reference's fp16_to_float on all 65 536 halves (float32 bit
    patterns, NaN payloads included), generated by tests/golden/make_golden.py."""
    want = np.load(os.path.join(HERE, "golden", "fp16_table.npz"))["bits"]
    got = O.keys_from_halfvec(np.arange(65536, dtype=np.uint16).reshape(1, -1))[0].view(np.uint32)
    assert np.array_equal(got, want)


def test_hnsw_heuristic_insert_mode():
    """Insert mode 3 (diversity heuristic + re-selection of full neighbours, an extension): a valid graph
    (no self links, no duplicates, counts within 2m) that answers at least as well as the reference rule."""
    n, dim, m = 6000, 16, 8
    X = W.mixture(n, dim, 16, 77)
    Q = W.mixture(200, dim, 16, 78, centers_seed=77)
    gt = W.exact_ground_truth(X, Q, 10)
    levels = O.hnsw_levels(n, seed=4)
    rec = {}
    for mode in (1, 3):
        g = O.Hnsw(dim, m, 32, 32, capacity=n)
        g.build(X, levels, mode)
        e = g.export()
        cnt0 = e["cnt"][:, 0]
        assert cnt0.max() <= 2 * m
        for v in range(0, n, 97):
            nb = e["nbr0"][v, :cnt0[v]]
            assert v not in nb and len(set(nb.tolist())) == len(nb) and np.all(nb < n)
        d, nn, _ = g.search(Q, 32, 10, 1, 1)
        rec[mode] = O.recall_at_k(nn.astype(np.int64), gt)
    assert rec[3] >= rec[1] - 0.005 and rec[3] >= 0.9, rec


# ---- the index access methods' own leaf functions (oracle/extract_ref_leafs.py) ------------------------
def _leaf_inputs():
    rng = np.random.default_rng(2718)
    out = []
    for dim in list(range(1, 40)) + [64, 128, 300, 768]:
        a = rng.standard_normal((3, dim)).astype(np.float32)
        b = rng.standard_normal((3, dim)).astype(np.float32)
        a[2] = 0.0                                   # zero norm: IVF cosine -> 1.0, HNSW cosine -> 2.0
        out.append((a, b))
    return out


@pytest.mark.skipif(O.ref_leafs_lib() is None, reason="oracle/_ref not built (reference tree absent)")
def test_index_distances_equal_the_reference_functions():
    """orc_ivf_distance / orc_hnsw_distance against ivfComputeDistance (ivf_am.c:1550-1592) and
    hnswComputeDistance (hnsw_am.c:1301-1345) themselves, compiled from the reference source."""
    ref = O.ref_leafs_lib()
    for a, b in _leaf_inputs():
        dim = a.shape[1]
        for i in range(a.shape[0]):
            for s in (1, 2):
                want = np.float32(ref.ref_ivf_distance(a[i], b[i], dim, s))
                got = O.distance_pairs(a[i:i + 1], b[i:i + 1], s, O.ARITH_IVF_F32)[0]
                assert BITS(np.array([got]))[0] == BITS(np.array([want]))[0], ("ivf", dim, s)
            for s in (1, 2, 3):
                want = np.float32(ref.ref_hnsw_distance(a[i], b[i], dim, s))
                got = O.distance_pairs(a[i:i + 1], b[i:i + 1], s, O.ARITH_HNSW)[0]
                assert BITS(np.array([got]))[0] == BITS(np.array([want]))[0], ("hnsw", dim, s)


@pytest.mark.skipif(O.ref_leafs_lib() is None, reason="oracle/_ref not built (reference tree absent)")
def test_kmeans_equals_the_reference_functions():
    """orc_kmeans_train against the reference's kmeans_init + kmeans_run (ivf_am.c:2070-2294), compiled
    from its source: centroids bit for bit, assignments, counts -- including k > n and empty clusters."""
    for n, dim, k in [(1200, 16, 12), (3000, 24, 40), (500, 7, 64), (30, 5, 50), (2000, 128, 100)]:
        X = W.mixture(n, dim, max(2, k // 3), n)
        rC, ra, rc = O.ref_kmeans_train(X, k)
        C, a, c, iters, cost = O.kmeans_train(X, k)
        assert np.array_equal(BITS(C), BITS(rC)) and np.array_equal(a, ra) and np.array_equal(c, rc), (n, dim, k)


@pytest.mark.skipif(O.ref_leafs_lib() is None, reason="oracle/_ref not built (reference tree absent)")
def test_level_draw_equals_the_reference_function():
    import ctypes
    libc = ctypes.CDLL(None)
    ref = O.ref_leafs_lib()
    libc.srandom(42)
    want = [ref.ref_hnsw_random_level(0.36) for _ in range(3000)]
    libc.srandom(42)
    got = [O.lib().orc_hnsw_random_level(0.36) for _ in range(3000)]
    assert got == want and max(want) >= 1


def test_index_leaf_golden_vectors():
    """The same comparisons against committed outputs of the reference's functions (tests/golden/
    index_leafs.npz, written by make_golden.py from oracle/_ref/libndb_ref_leafs.so), so that they also
    run where the reference tree is absent."""
    g = np.load(os.path.join(HERE, "golden", "index_leafs.npz"))
    ivf, hn = [], []
    for a, b in _leaf_inputs():
        for s in (1, 2):
            ivf.append(O.distance_pairs(a, b, s, O.ARITH_IVF_F32))
        for s in (1, 2, 3):
            hn.append(O.distance_pairs(a, b, s, O.ARITH_HNSW))
    assert np.array_equal(BITS(np.concatenate(ivf)), g["ivf_bits"])
    assert np.array_equal(BITS(np.concatenate(hn)), g["hnsw_bits"])
    X = W.mixture(1500, 16, 10, 31)
    C, a, c, _, _ = O.kmeans_train(X, 24)
    assert np.array_equal(BITS(C), g["km_C_bits"]) and np.array_equal(a, g["km_assign"]) and np.array_equal(c, g["km_counts"])


# ---- the reference's index fixture (t/010_indexes_comprehensive.t:32-48) ------------------------------
FIXTURE_T010 = np.array([[1 + i, 2 + i, 3 + i, 4 + i] for i in range(8)], np.float32)


def test_reference_index_fixture_t010():
    """The 8-row table of the reference's index test, query '[1,2,3,4]' ORDER BY vec <-> q: the reference
    asserts no output, but the answer is forced -- row i is at distance exactly 2 i."""
    X, q = FIXTURE_T010, FIXTURE_T010[:1]
    want_d = np.arange(8, dtype=np.float32) * 2
    d, i = O.knn_exact(X, q, 8, 1, O.ARITH_OP_F64)
    assert np.array_equal(i[0], np.arange(8)) and np.array_equal(d[0], want_d)
    C, assign, counts, _, _ = O.kmeans_train(X, 4)
    lists = O.ivf_assign(X, C)
    off, rows = O.lists_from_assignment(lists, 4)
    d, i, _ = O.ivf_search(X, C, off, rows, q, 4, 8)
    assert np.array_equal(i[0], np.arange(8)) and np.array_equal(d[0], want_d)
    g = O.Hnsw(4, 16, 200, 64, capacity=8)
    g.build(X, np.zeros(8, np.int32), 1)
    d, n, _ = g.search(q, 64, 8, 1, 1)
    assert np.array_equal(n[0], np.arange(8)) and np.array_equal(d[0], want_d)


# ---- knn_classify / knn_regress / cluster_kmeans (oracle/ndb_oracle_ml.c) -------------------
def _ml_cases():
    return [(400, 8, 5, 3), (900, 33, 4, 7), (1500, 16, 12, 5), (60, 3, 6, 60)]     # n, dim, k (clusters / neighbours), seed


@pytest.mark.skipif(O.ref_leafs_lib() is None, reason="oracle/_ref not built (reference tree absent)")
def test_ml_distances_equal_the_reference_functions():
    """orc_ml_euclidean against euclidean_distance (ml_knn.c:76-90) and orc_l2_distance_squared against
    neurondb_l2_distance_squared (neurondb_simd_impl.c:36-104), both compiled from the reference source."""
    ref = O.ref_leafs_lib()
    for a, b in _leaf_inputs():
        for i in range(a.shape[0]):
            dim = a.shape[1]
            assert O.lib().orc_ml_euclidean(a[i], b[i], dim) == ref.ref_ml_euclidean(a[i], b[i], dim)
            assert O.lib().orc_l2_distance_squared(a[i], b[i], dim) == ref.ref_l2_distance_squared(a[i], b[i], dim)


@pytest.mark.skipif(O.ref_leafs_lib() is None, reason="oracle/_ref not built (reference tree absent)")
def test_knn_ml_equals_the_reference_functions():
    for n, dim, k, seed in _ml_cases():
        rng = np.random.default_rng(seed)
        X = W.mixture(n, dim, 5, seed)
        Q = W.mixture(30, dim, 5, seed + 1, centers_seed=seed)
        labels = rng.integers(0, 3, n).astype(np.float64) + rng.integers(0, 2, n) * 0.5
        cls, mean, rows = O.knn_ml(X, labels, Q, min(k, n))
        rcls, rmean, rdist = O.ref_knn_ml(X, labels, Q, min(k, n))
        assert np.array_equal(cls, rcls) and np.array_equal(mean, rmean)
        for j in range(len(Q)):                                      # the rows are the k smallest reference distances
            assert np.array_equal(np.sort(rdist[j])[:rows.shape[1]], rdist[j][rows[j]])


@pytest.mark.skipif(O.ref_leafs_lib() is None, reason="oracle/_ref not built (reference tree absent)")
def test_cluster_kmeans_equals_the_reference_functions():
    """orc_cluster_kmeans against the reference's kmeanspp_init (ml_kmeans.c:45-139) and the text of cluster_kmeans'
    Lloyd loop (:226-278) compiled from its source, with rand() seeded alike: labels, centers bit for bit, iterations."""
    import ctypes
    for n, dim, k, seed in _ml_cases():
        X = W.mixture(n, dim, max(2, k // 2), seed)
        if n == 60:
            X[10:20] = X[0]                                           # duplicates: zero D^2 weights, empty clusters
        for max_iters in (0, 3):
            draws = O.libc_rand_draws(seed, k)
            labels, centers, seeds, it = O.cluster_kmeans(X, k, max_iters, draws)
            rl, rc, rit = O.ref_cluster_kmeans(X, k, max_iters, seed)
            assert it == rit and np.array_equal(labels, rl) and np.array_equal(BITS(centers), BITS(rc)), (n, dim, k)
            ctypes.CDLL(None).srand(seed)
            rs = np.zeros(k, np.int32)
            O.ref_leafs_lib().ref_kmeanspp_init(X, n, dim, k, rs)
            assert np.array_equal(seeds, rs)


def test_ml_golden_vectors():
    """The same against tests/golden/ml_paths.npz (outputs of the reference's functions, written by make_golden.py)."""
    g = np.load(os.path.join(HERE, "golden", "ml_paths.npz"))
    for n, dim, k, seed in _ml_cases():
        tag = "n%d" % n
        X = W.mixture(n, dim, max(2, k // 2), seed)
        labels, centers, seeds, it = O.cluster_kmeans(X, k, 0, g["draws_" + tag])
        assert np.array_equal(labels, g["km_labels_" + tag]) and np.array_equal(BITS(centers), g["km_center_bits_" + tag])
        assert it == int(g["km_iters_" + tag])
        rng = np.random.default_rng(seed)
        Q = W.mixture(30, dim, 5, seed + 1, centers_seed=seed)
        lab = rng.integers(0, 3, n).astype(np.float64)
        cls, mean, _ = O.knn_ml(X, lab, Q, min(k, n))
        assert np.array_equal(cls, g["knn_cls_" + tag]) and np.array_equal(mean, g["knn_mean_" + tag])


# ---- product quantisation (oracle/ndb_oracle_ml.c) --------------------------------------------
def _pq_cases():
    return [(600, 16, 4, 16, 11), (900, 24, 8, 32, 12), (300, 12, 3, 256, 13), (50, 8, 8, 2, 14)]     # n, dim, m, ksub, seed


@pytest.mark.skipif(O.ref_leafs_lib() is None, reason="oracle/_ref not built (reference tree absent)")
def test_pq_equals_the_reference_functions():
    """orc_pq_train / _encode / _asymmetric_distance against the reference's train_subspace_kmeans (ml_product_
    quantization.c:80-190) and the text of the loops of pq_encode_vector (:479-503) and pq_asymmetric_distance
    (:1063-1098), compiled from its source; rand() seeded alike."""
    for n, dim, m, ksub, seed in _pq_cases():
        X = W.mixture(n, dim, 6, seed)
        Q = W.mixture(12, dim, 6, seed + 1, centers_seed=seed)
        draws = O.libc_rand_draws(seed, m * ksub)
        cb = O.pq_train(X, m, ksub, draws, 5)
        assert np.array_equal(BITS(cb), BITS(O.ref_pq_train(X, m, ksub, seed, 5))), (n, dim, m, ksub)
        codes = O.pq_encode(X, cb)
        assert np.array_equal(codes, O.ref_pq_encode(X, cb))
        d, r, al = O.pq_knn(Q, codes, cb, 5, want_all=True)
        assert np.array_equal(BITS(al), BITS(O.ref_pq_distances(Q, codes, cb)))


def test_pq_golden_vectors():
    g = np.load(os.path.join(HERE, "golden", "ml_paths.npz"))
    for n, dim, m, ksub, seed in _pq_cases():
        tag = "pq%d" % n
        X = W.mixture(n, dim, 6, seed)
        Q = W.mixture(12, dim, 6, seed + 1, centers_seed=seed)
        cb = O.pq_train(X, m, ksub, g["draws_" + tag], 5)
        assert np.array_equal(BITS(cb), g["cb_bits_" + tag])
        codes = O.pq_encode(X, cb)
        assert np.array_equal(codes, g["codes_" + tag])
        _, _, al = O.pq_knn(Q, codes, cb, 5, want_all=True)
        assert np.array_equal(BITS(al), g["adc_bits_" + tag])


# ---- per-vector quantisers (src/types/quantization.c) ------------------------------------------
def _quant_inputs():
    rng = np.random.default_rng(4242)
    out = []
    for dim in (1, 2, 3, 7, 8, 9, 16, 31, 64, 100, 128):
        X = (rng.standard_normal((12, dim)) * rng.choice([1e-3, 1.0, 300.0], (12, 1))).astype(np.float32)
        X[0] = 0.0                                            # all zeros: zero bytes in every format
        X[1] = X[1, 0]                                        # constant row: uint8's max == min
        X[2, ::2] = 0.0
        X[3] *= 1e-6                                          # below fp16's normal range: flushed to zero
        X[4] *= 1e6                                           # above it: inf
        X[5] = np.round(X[5] * 4) / 4 + 0.5                   # .5 ties for rintf
        out.append(X)
    return out


@pytest.mark.skipif(O.ref_leafs_lib() is None, reason="oracle/_ref not built (reference tree absent)")
def test_quantisers_equal_the_reference_functions():
    ref = O.ref_leafs_lib()
    for X in _quant_inputs():
        for kind in (O.Q_INT8, O.Q_FP16, O.Q_BINARY, O.Q_UINT8, O.Q_TERNARY, O.Q_INT4):
            assert np.array_equal(O.quantize_rows(kind, X), O.ref_quantize_rows(kind, X)), (kind, X.shape)
        bits = O.quantize_rows(O.Q_BINARY, X)
        for i in range(len(bits)):
            j = (i + 1) % len(bits)
            want = ref.ref_hamming(bits[i], bits[j], X.shape[1])
            assert O.lib().orc_hamming(bits[i], bits[j], X.shape[1]) == want == int(np.unpackbits(bits[i] ^ bits[j]).sum())


def test_quantiser_golden_vectors():
    g = np.load(os.path.join(HERE, "golden", "ml_paths.npz"))
    for X in _quant_inputs():
        for kind in (O.Q_INT8, O.Q_FP16, O.Q_BINARY, O.Q_UINT8, O.Q_TERNARY, O.Q_INT4):
            assert np.array_equal(O.quantize_rows(kind, X), g["quant_k%d_d%d" % (kind, X.shape[1])])


# ---- cluster_minibatch_kmeans (ml_minibatch_kmeans.c) ------------------------------------------
def _minibatch_cases():
    return [(800, 8, 5, 50, 20, 21), (1500, 24, 12, 100, 30, 22), (300, 6, 7, 1000, 5, 23), (40, 3, 8, 16, 12, 24)]   # n, dim, k, batch, iters, seed


def _minibatch_rows(n, dim, k, seed):
    X = W.mixture(n, dim, max(2, k // 2), seed)
    if n == 40:
        X[:] = X[:4].repeat(10, axis=0)          # 4 distinct rows for 8 clusters: the seeding stops early (sum < 1e-10)
    return X


@pytest.mark.skipif(O.ref_leafs_lib() is None, reason="oracle/_ref not built (reference tree absent)")
def test_cluster_minibatch_kmeans_equals_the_reference_functions():
    """orc_cluster_minibatch_kmeans against the reference's minibatch_kmeans_pp_init (ml_minibatch_kmeans.c:67-198) and the
    text of cluster_minibatch_kmeans' main loop and final assignment (:347-425), rand() seeded alike."""
    for n, dim, k, batch, iters, seed in _minibatch_cases():
        X = _minibatch_rows(n, dim, k, seed)
        draws = O.libc_rand_draws(seed, k + batch * iters)
        labels, centers, used = O.cluster_minibatch_kmeans(X, k, batch, iters, draws)
        rl, rc = O.ref_cluster_minibatch_kmeans(X, k, batch, iters, seed)
        assert np.array_equal(labels, rl) and np.array_equal(BITS(centers), BITS(rc)), (n, dim, k)
        assert used == (k if n != 40 else 4) + min(batch, n) * iters


def test_cluster_minibatch_kmeans_golden_vectors():
    g = np.load(os.path.join(HERE, "golden", "ml_paths.npz"))
    for n, dim, k, batch, iters, seed in _minibatch_cases():
        X = _minibatch_rows(n, dim, k, seed)
        labels, centers, used = O.cluster_minibatch_kmeans(X, k, batch, iters, g["mb_draws_n%d" % n])
        assert np.array_equal(labels, g["mb_labels_n%d" % n]) and np.array_equal(BITS(centers), g["mb_center_bits_n%d" % n])


def test_quantisation_fixture_of_the_reference_sql_tests():
    """sql/11_quantization_detail.sql quantises '[1,-1,0,3]' with every method and states the sizes ("4 dimensions use 1 byte"
    for binary and ternary, 2 bytes for int4, 1 / 2 bytes per dimension for int8 / uint8 / fp16); the bytes follow from the
    functions' definitions."""
    v = np.array([[1, -1, 0, 3]], np.float32)
    sizes = {O.Q_INT8: 4, O.Q_FP16: 8, O.Q_BINARY: 1, O.Q_UINT8: 4, O.Q_TERNARY: 1, O.Q_INT4: 2}
    for kind, nbytes in sizes.items():
        assert O.lib().orc_quantized_row_bytes(kind, 4) == nbytes
    assert O.quantize_rows(O.Q_BINARY, v)[0].tolist() == [0b1001]                       # bits 0 and 3: the components > 0
    assert O.quantize_rows(O.Q_INT8, v)[0].view(np.int8).tolist() == [42, -42, 0, 127]  # rintf(x * 127 / 3)
    assert O.quantize_rows(O.Q_UINT8, v)[0].tolist() == [128, 0, 64, 255]               # rintf((x + 1) * 255 / 4): 127.5 -> 128, 63.75 -> 64
    assert O.quantize_rows(O.Q_TERNARY, v)[0].tolist() == [0b10000000]                  # threshold 3/3 = 1, strict: 1 and -1 are 0; 3 -> 2 in bits 6-7
    assert O.quantize_rows(O.Q_FP16, v)[0].view(np.uint16).tolist() == [0x3C00, 0xBC00, 0x0000, 0x4200]
    assert O.quantize_rows(O.Q_INT4, v)[0].tolist() == [(8 - 2) << 4 | (8 + 2), (8 + 7) << 4 | 8]
