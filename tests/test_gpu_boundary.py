"""GPU: the drop-in boundary exercised the way the reference reaches it -- through the ndb_gpu_backend struct
(surface b2) and through the index-AM scan / staging shim (surface b1), both compiled from
integration/*.c against the reference's own header."""
import ctypes as C
import os

import numpy as np
import pytest

import workloads as W

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BITS = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)


class Tid(C.Structure):
    _fields_ = [("bi_hi", C.c_uint16), ("bi_lo", C.c_uint16), ("ip_posid", C.c_uint16)]


@pytest.fixture(scope="module")
def glue(ndb):
    path = os.path.join(ROOT, "oracle", "_ref", "libndb_b200_glue.so")
    assert os.path.exists(path), "oracle/_ref/libndb_b200_glue.so must travel with the snapshot (make glue)"
    g = C.CDLL(path)
    g.ndb_b200_am_ivf_beginscan.restype = C.c_void_p
    g.ndb_b200_am_ivf_beginscan.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    g.ndb_b200_am_hnsw_beginscan.restype = C.c_void_p
    g.ndb_b200_am_hnsw_beginscan.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    g.ndb_b200_am_rescan.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    g.ndb_b200_am_gettuple.argtypes = [C.c_void_p, C.POINTER(Tid), C.POINTER(C.c_float)]
    g.ndb_b200_am_endscan.argtypes = [C.c_void_p]
    g.ndb_b200_am_endscan.restype = None
    g.ndb_b200_am_stage_ivf.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    g.ndb_b200_am_stage_hnsw.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    assert g.ndb_b200_glue_init() == 0
    return g


def test_launchers_through_the_backend_struct(ndb, orc, glue):
    """t/005's known answers and seeded pairs through backend->launch_l2_distance / ->launch_cosine."""
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    A = np.array([[0, 0, 0], [1, 2, 3], [1, 0, 0]], np.float32)
    B = np.array([[3, 4, 0], [1, 2, 3], [0, 1, 0]], np.float32)
    out = np.zeros(3, np.float32)
    assert glue.ndb_b200_glue_l2(p(A), p(B), p(out), 3, 3) == 0
    assert out[0] == 5.0 and out[1] == 0.0
    assert glue.ndb_b200_glue_cosine(p(A), p(B), p(out), 3, 3) == 0
    assert out[2] == 1.0                                  # orthogonal
    A, B = W.gaussian(1000, 96, 1), W.gaussian(1000, 96, 2)
    out = np.zeros(1000, np.float32)
    assert glue.ndb_b200_glue_l2(p(A), p(B), p(out), 1000, 96) == 0
    assert np.array_equal(BITS(out), BITS(ndb.launch_l2_distance(A, B)))
    want = np.sqrt(((A.astype(np.float64) - B) ** 2).sum(1))
    assert np.max(np.abs(out - want) / want) < 1e-5
    # k-means step through the struct == the oracle's kmeans_assign / update_centroids
    X, Cn = W.mixture(3000, 16, 8, 5), W.mixture(8, 16, 8, 6)
    idx = np.zeros(3000, np.int32)
    assert glue.ndb_b200_glue_kmeans_assign(p(X), p(Cn), p(idx), 3000, 16, 8) == 0
    assert np.array_equal(idx, orc.kmeans_assign(X, Cn))
    C2 = np.zeros((8, 16), np.float32)
    assert glue.ndb_b200_glue_kmeans_update(p(X), p(idx), p(C2), 3000, 16, 8) == 0
    assert np.array_equal(BITS(C2), BITS(orc.kmeans_update(X, idx, 8)[0]))
    # launch_pq_encode through the struct == pq_encode_vector's loop (byte codes)
    cb = W.gaussian(4 * 16, 4, 7).reshape(4, 16, 4)
    codes = np.zeros((3000, 4), np.uint8)
    assert glue.ndb_b200_glue_pq_encode(p(X), p(cb), p(codes), 3000, 16, 4, 16) == 0
    assert np.array_equal(codes.astype(np.int16), orc.pq_encode(X, cb))
    assert glue.ndb_b200_glue_stream_roundtrip() == 0
    name = C.create_string_buffer(256)
    total, major = C.c_size_t(), C.c_int()
    assert glue.ndb_b200_glue_device_info(0, name, 256, C.byref(total), C.byref(major)) == 0
    assert b"B200" in name.value and total.value > 100 << 30 and major.value == 10


def tid(i):
    return ((i // 100 + 1) << 16) | (i % 100 + 1)


def _scan(glue, so, q, k):
    assert glue.ndb_b200_am_rescan(so, q.ctypes.data_as(C.c_void_p), q.shape[0], k) == 0
    t, d = Tid(), C.c_float()
    tids, dist = [], []
    while True:
        rc = glue.ndb_b200_am_gettuple(so, C.byref(t), C.byref(d))
        assert rc >= 0
        if rc == 0:
            break
        tids.append((((t.bi_hi << 16) | t.bi_lo) << 16) | t.ip_posid)
        dist.append(d.value)
    assert glue.ndb_b200_am_gettuple(so, C.byref(t), C.byref(d)) == 0       # exhausted stays exhausted
    return np.array(tids, np.int64), np.array(dist, np.float32)


@pytest.mark.parametrize("mode", ["literal", "full"])
def test_ivf_scan_state_machine_over_a_staged_relation(ndb, orc, glue, mode):
    """ambeginscan / amrescan / amgettuple / amendscan over an index staged from its page image: the TIDs and
    ORDER BY distances ivfgettuple would hand the executor, equal to the oracle's ivfSelectClusters +
    ivfCollectCandidates on the same pages (literal = the k*10 candidate cut-off of ivf_am.c:1743)."""
    n, dim, lists, nprobe, k = 5000, 24, 16, 4, 10
    X = W.mixture(n, dim, lists, 91)
    Q = W.mixture(30, dim, lists, 92, centers_seed=91)
    tids = np.array([tid(i) for i in range(n)], np.int64)
    Cn, _, _, _, _ = orc.kmeans_train(X[:lists * 100], lists)
    assign = orc.ivf_assign(X, Cn)
    nb, blocks = orc.ivf_encode_relation(X, Cn, assign, tids)
    blocks = np.ascontiguousarray(blocks).reshape(-1)
    ix = ndb.IvfIndex(dim, lists)
    reader = C.cast(glue.ndb_b200_am_image_reader, C.c_void_p)
    assert glue.ndb_b200_am_stage_ivf(ix.h, reader, None, blocks.ctypes.data_as(C.c_void_p), blocks.size // 8192) == 0
    assert len(ix) == n
    off, rows = orc.lists_from_assignment(assign, lists)
    literal = mode == "literal"
    so = glue.ndb_b200_am_ivf_beginscan(ix.h, nprobe, ndb.IVF_LITERAL if literal else ndb.IVF_FULL, ndb.ARITH_IVF_F32)
    assert so
    od, oi, _ = orc.ivf_search(X, Cn, off, rows, Q, nprobe, k, literal=literal, ids=tids)
    for qi in range(Q.shape[0]):
        got_t, got_d = _scan(glue, so, Q[qi], k)
        want = oi[qi] >= 0
        assert np.array_equal(got_t, oi[qi][want]) and np.array_equal(BITS(got_d), BITS(od[qi][want]))
        assert np.all(np.diff(got_d) >= 0)
    # a vector of another dimension is the reference's dimension error; no query = no tuples
    bad = np.zeros(dim + 1, np.float32)
    assert glue.ndb_b200_am_rescan(so, bad.ctypes.data_as(C.c_void_p), dim + 1, k) == -5
    glue.ndb_b200_am_endscan(so)
    so2 = glue.ndb_b200_am_ivf_beginscan(ix.h, nprobe, ndb.IVF_FULL, ndb.ARITH_IVF_F32)
    assert glue.ndb_b200_am_gettuple(so2, C.byref(Tid()), None) == 0
    glue.ndb_b200_am_endscan(so2)


def test_hnsw_scan_state_machine_over_a_staged_relation(ndb, orc, glue):
    n, dim, m, ef, k = 800, 20, 8, 32, 10
    X = W.gaussian(n, dim, 17)
    Q = W.gaussian(20, dim, 18)
    tids = np.array([tid(i) for i in range(n)], np.int64)
    levels = orc.hnsw_levels(n, seed=5)
    g = orc.Hnsw(dim, m, 32, ef, capacity=n)
    g.build(X, levels, 1)
    nb, blocks = orc.hnsw_encode_relation(g, X, tids)
    blocks = np.ascontiguousarray(blocks).reshape(-1)
    h = ndb.HnswIndex(dim, m, 32, ef)
    reader = C.cast(glue.ndb_b200_am_image_reader, C.c_void_p)
    assert glue.ndb_b200_am_stage_hnsw(h.h, reader, None, blocks.ctypes.data_as(C.c_void_p), blocks.size // 8192) == 0
    for mode, smode in ((ndb.HNSW_LITERAL, 0), (ndb.HNSW_BESTFIRST, 1)):
        so = glue.ndb_b200_am_hnsw_beginscan(h.h, dim, ef, mode)
        od, on, _ = g.search(Q, ef, k, 1, smode)
        for qi in range(Q.shape[0]):
            got_t, got_d = _scan(glue, so, Q[qi], k)
            want = on[qi] != 0xFFFFFFFF
            assert np.array_equal(got_t, tids[on[qi][want].astype(np.int64)]), (mode, qi)
            assert np.array_equal(BITS(got_d), BITS(od[qi][want]))
        glue.ndb_b200_am_endscan(so)

