"""GPU: the tcgen05 (bf16 tensor-core) GEMM-form distance path, tolerance 1e-3 (north_star)."""
import ctypes as C

import numpy as np
import pytest

import workloads as W

pytestmark = pytest.mark.gpu


def bf16_round(a):
    """Round fp32 to the nearest bf16-representable value (round to nearest even)."""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


@pytest.mark.parametrize("dim", [128, 64, 96, 200, 256, 384, 768, 1000])      # > 256: query tile streamed
def test_umma_tile_matches_numpy(ndb, dim):
    """Raw accumulator D = Q X^T of one 128 x 256 tile: validates the smem/instruction descriptors."""
    rng = np.random.default_rng(dim)
    Q = bf16_round(rng.standard_normal((128, dim)).astype(np.float32))
    X = bf16_round(rng.standard_normal((256, dim)).astype(np.float32))
    D = np.zeros((128, 256), np.float32)
    lib = ndb._lib.load()
    lib.ndbdbg_tc_gemm.restype = C.c_int
    lib.ndbdbg_tc_gemm.argtypes = [C.c_void_p] * 1 + [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    rc = lib.ndbdbg_tc_gemm(Q.ctypes.data, 128, X.ctypes.data, 256, dim, D.ctypes.data)
    assert rc == 0, lib.ndb_b200_last_error()
    want = Q.astype(np.float64) @ X.astype(np.float64).T
    assert np.max(np.abs(D - want)) < 1e-3 * np.max(np.abs(want)), np.max(np.abs(D - want))


@pytest.mark.parametrize("n,dim,nq,k", [(5000, 128, 300, 10), (1000, 96, 130, 1), (70000, 128, 1000, 10), (300, 256, 7, 16),
                                         (6000, 768, 300, 10), (3000, 1536, 40, 10)])
@pytest.mark.parametrize("metric", [1, 2, 3])
def test_tensor_knn_within_tolerance(ndb, orc, n, dim, nq, k, metric):
    X = bf16_round(W.gaussian(n, dim, 50 + n))          # config 5 data is bf16: inputs are representable
    Q = bf16_round(W.gaussian(nq, dim, 51 + n))
    ds = ndb.Dataset(dim)
    ds.append(X)
    d, i = ds.knn(Q, k, metric, ndb.ARITH_TENSOR)
    if metric in (1, 2):
        od, oi = orc.knn_exact(X, Q, k, metric, orc.ARITH_OP_F64)
    else:
        # tensor IP = -dot (ranking by largest inner product, hnsw_am.c:1334-1337 sign convention)
        dots = Q.astype(np.float64) @ X.astype(np.float64).T
        oi = np.argsort(-dots, axis=1, kind="stable")[:, :k]
        od = np.take_along_axis(-dots, oi, 1).astype(np.float32)
    # distances within 1e-3 relative (bf16 tensor-core contract)
    assert np.max(np.abs(d - od) / np.maximum(np.abs(od), 1e-3)) < 1e-3
    # ids identical wherever the gap to the neighbouring rank exceeds the tolerance
    same = (i == oi)
    assert same.mean() > 0.995
    for q, j in zip(*np.nonzero(~same)):
        assert abs(od[q, j] - d[q, j]) <= 1e-3 * max(abs(od[q, j]), 1e-3)
    assert np.all(np.diff(d, axis=1) >= 0)


IVF_CASES = [
    (20000, 128, 64, 700, 8, 10, 1),      # several query tiles per list
    (5000, 96, 40, 33, 16, 10, 3),        # ragged dims, IP
    (3000, 128, 100, 200, 100, 5, 1),     # nprobe > 32: exact fp32 coarse stage, all lists probed
    (150000, 64, 16, 300, 4, 16, 1),      # long lists -> several segments per list
    (400, 256, 8, 5, 3, 10, 1),           # two K-chunks, tiny lists
    (8000, 384, 32, 300, 8, 10, 1),       # three K-chunks: the query tile is streamed with the stored tiles
    (12000, 64, 48, 400, 8, 10, 2),       # vector_cosine_ops: lists scanned by cosine distance (coarse stage stays L2)
    (2000, 128, 16, 50, 20, 10, 2),       # cosine, 64-candidate coarse certificate (nprobe in 17..32), short lists
    (300, 64, 32, 50, 16, 10, 1),         # lists shorter than the candidate count: pad rows must never rank
    (40000, 96, 512, 600, 32, 10, 3),     # C4-shaped: inner product, nprobe = 32 on the tensor coarse stage
    (30000, 128, 128, 500, 16, 1, 1),     # k = 1
    (60000, 64, 64, 300, 8, 24, 1),       # k > 16: 16-entry partial lists, 64 candidates re-evaluated per query
    (25000, 96, 48, 200, 12, 32, 3),      # k = 32, inner product
    (20000, 32, 32, 150, 6, 32, 2),       # k = 32, cosine
]


def _assert_equals_fp32_path(ndb, ix, Q, nprobe, k):
    d0, i0 = ix.search(Q, nprobe, k, ndb.IVF_FULL, ndb.ARITH_IVF_F32)
    d1, i1 = ix.search(Q, nprobe, k, ndb.IVF_FULL, ndb.ARITH_TENSOR)
    st = ix.cert_stats()
    assert np.array_equal(i1, i0), ("ids differ on %d of %d queries" % ((i1 != i0).any(1).sum(), Q.shape[0]), st)
    assert np.array_equal(d1.view(np.uint32), d0.view(np.uint32)), st
    return st


@pytest.mark.parametrize("n,dim,lists,nq,nprobe,k,metric", IVF_CASES)
@pytest.mark.parametrize("representable", [False, True])
def test_tensor_ivf_equals_fp32_path(ndb, orc, n, dim, lists, nq, nprobe, k, metric, representable):
    """arith=TENSOR proposes candidates with bf16 products, re-evaluates them in the reference's fp32 arithmetic and
    certifies the answer (or recomputes it exactly): ids AND distance bits equal the fp32 path's -- on bf16-valued
    inputs (config 5's) and on ordinary fp32 mixtures, whose bf16 rounding moves every key."""
    X = W.mixture(n, dim, max(lists // 2, 2), 900 + n)
    Q = W.mixture(nq, dim, max(lists // 2, 2), 977 + n, centers_seed=900 + n)
    if representable:
        X, Q = bf16_round(X), bf16_round(Q)
    ix = ndb.IvfIndex(dim, lists, metric)
    ix.ivfbuild(X)
    ix.ivfinsert(X)
    st = _assert_equals_fp32_path(ndb, ix, Q, nprobe, k)
    # the certificate, not the exact kernel, must carry clustered data
    assert st["list_full_scan_queries"] <= 0.02 * nq + 2, st
    if lists >= 256:           # (a centroid store of less than a tile leaves the coarse certificate no margin)
        assert st["coarse_fallback_queries"] <= 0.05 * nq + 2, st


@pytest.mark.parametrize("n,dim,lists,nq,nprobe,k,metric", IVF_CASES)
def test_tensor_ivf_two_phase_scan_equals_fp32_path(ndb, monkeypatch, n, dim, lists, nq, nprobe, k, metric):
    """The two-phase scan (first segment of every query's nearest list, the bound it leaves, then the rest) is chosen
    by the library for long lists only; forced here on every shape -- lists of one segment and of several, replicated
    query tiles, every metric -- and held to the same contract: ids and distance bits of the fp32 path."""
    monkeypatch.setenv("NDB_IVF_TC_PHASES", "2")
    X = W.mixture(n, dim, max(lists // 2, 2), 900 + n)
    Q = W.mixture(nq, dim, max(lists // 2, 2), 977 + n, centers_seed=900 + n)
    ix = ndb.IvfIndex(dim, lists, metric)
    ix.ivfbuild(X)
    ix.ivfinsert(X)
    _assert_equals_fp32_path(ndb, ix, Q, nprobe, k)
    monkeypatch.setenv("NDB_IVF_TC_PHASES", "1")
    _assert_equals_fp32_path(ndb, ix, Q, nprobe, k)


@pytest.mark.parametrize("metric", [1, 2, 3])
def test_tensor_ivf_exact_on_structureless_data(ndb, metric):
    """Isotropic Gaussian rows: neighbours are nearly equidistant, the rounding bound often cannot separate the k-th
    from the next candidates, and those queries go through the exact kernel.  The result must not care."""
    X = W.gaussian(30000, 64, 31)
    Q = W.gaussian(400, 64, 32)
    Q[7] = 0.0                               # a zero query: cosine distance is 1.0f to everything, ties by id
    Q[11] = X[123]                           # an exact hit
    ix = ndb.IvfIndex(64, 64, metric)
    ix.ivfbuild(X)
    ix.ivfinsert(X)
    _assert_equals_fp32_path(ndb, ix, Q, 8, 10)
    _assert_equals_fp32_path(ndb, ix, Q, 24, 16)


def test_tensor_ivf_equals_oracle_c2_slice(ndb, orc):
    """A C2-shaped slice (L2, 128-d mixture, non-representable fp32) against the ORACLE: ids and distance bits."""
    n, dim, lists, nq, nprobe, k = 100_000, 128, 128, 400, 16, 10
    X = W.mixture(n, dim, 128, 2024)
    Q = W.mixture(nq, dim, 128, 2025, centers_seed=2024)
    ix = ndb.IvfIndex(dim, lists)
    ix.ivfbuild(X)
    got_lists = ix.ivfinsert(X)
    off, rows = orc.lists_from_assignment(got_lists, lists)
    d, i = ix.search(Q, nprobe, k, ndb.IVF_FULL, ndb.ARITH_TENSOR)
    od, oi, _ = orc.ivf_search(X, ix.centroids(), off, rows, Q, nprobe, k)
    assert np.array_equal(i, oi)
    assert np.array_equal(d.view(np.uint32), od.view(np.uint32))


def test_tensor_ivf_rejects_what_it_cannot_do(ndb):
    X = W.gaussian(500, 32, 1)
    ix2 = ndb.IvfIndex(32, 4, ndb.L2)
    ix2.ivfbuild(X)
    ix2.ivfinsert(X)
    with pytest.raises(ndb.NdbError):
        ix2.search(X[:3], 2, 33, ndb.IVF_FULL, ndb.ARITH_TENSOR)      # k > 32
    with pytest.raises(ndb.NdbError):
        ix2.search(X[:3], 2, 10, ndb.IVF_LITERAL, ndb.ARITH_TENSOR)   # literal mode


def test_tensor_ivf_on_a_list_shard(ndb, orc):
    """One rank of a two-rank job (lists l % 2 == 1 only): most probed lists are empty here, the rest
    short -- the tile padding must stay invisible."""
    X = bf16_round(W.mixture(30000, 128, 128, 4242))
    Q = bf16_round(W.mixture(500, 128, 128, 4243, centers_seed=4242))
    ix = ndb.IvfIndex(128, 256)
    ix.set_shard(1, 2)
    ix.ivfbuild(X)
    ix.ivfinsert(X)
    d0, i0 = ix.search(Q, 20, 10, ndb.IVF_FULL, ndb.ARITH_IVF_F32)
    d1, i1 = ix.search(Q, 20, 10, ndb.IVF_FULL, ndb.ARITH_TENSOR)
    assert np.array_equal(i1, i0)
    assert np.array_equal(d1.view(np.uint32), d0.view(np.uint32))
