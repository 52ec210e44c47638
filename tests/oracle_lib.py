"""ctypes bindings for the CPU oracle (oracle/libndb_oracle.so) and, when present, the
reference's own operator sources compiled under oracle/pgshim (oracle/_ref/*.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (neurondb_b200/) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

L2, COSINE, IP = 1, 2, 3
ARITH_OP_F64, ARITH_AVX2, ARITH_AVX512, ARITH_IVF_F32, ARITH_HNSW = 0, 1, 2, 3, 4

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
Q_INT8, Q_FP16, Q_BINARY, Q_UINT8, Q_TERNARY, Q_INT4 = 1, 2, 3, 4, 5, 6
RAND_MAX = 2147483647                      # glibc


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)


def _load(name="libndb_oracle.so"):
    path = os.path.join(ORACLE_DIR, name)
    if not os.path.exists(path):
        build_oracle()
    lib = C.CDLL(path)
    f = lib.orc_check_vector; f.restype = C.c_int; f.argtypes = [_f32p, C.c_int]
    for nm in ("orc_l2_distance", "orc_inner_product_distance", "orc_inner_product_op",
               "orc_cosine_distance", "orc_kmeans_l2sq"):
        f = getattr(lib, nm); f.restype = C.c_float; f.argtypes = [_f32p, _f32p, C.c_int]
    for nm in ("orc_l2_avx", "orc_ip_avx", "orc_cosine_avx", "orc_ivf_distance", "orc_hnsw_distance"):
        f = getattr(lib, nm); f.restype = C.c_float; f.argtypes = [_f32p, _f32p, C.c_int, C.c_int]
    f = lib.orc_distance; f.restype = C.c_float; f.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int]
    f = lib.orc_distance_pairs; f.restype = None
    f.argtypes = [_f32p, _f32p, _f32p, C.c_int64, C.c_int, C.c_int, C.c_int]
    f = lib.orc_knn_exact; f.restype = None
    f.argtypes = [_f32p, C.c_void_p, C.c_int64, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int,
                  _f32p, _i64p, C.c_int]
    f = lib.orc_kmeans_train; f.restype = C.c_int
    f.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _f32p, _i32p, _i32p,
                  C.POINTER(C.c_float)]
    f = lib.orc_kmeans_assign; f.restype = None
    f.argtypes = [_f32p, C.c_int64, C.c_int, _f32p, C.c_int, _i32p, C.c_int]
    f = lib.orc_kmeans_update; f.restype = None
    f.argtypes = [_f32p, _i32p, C.c_int64, C.c_int, C.c_int, _f32p, _i32p]
    f = lib.orc_ivf_train_samples; f.restype = C.c_int; f.argtypes = [C.c_int64, C.c_int]
    f = lib.orc_ivf_assign; f.restype = None
    f.argtypes = [_f32p, C.c_int64, C.c_int, _f32p, C.c_int, _i32p, C.c_int]
    f = lib.orc_ivf_select_clusters; f.restype = None
    f.argtypes = [_f32p, C.c_int, _f32p, C.c_int, C.c_int, _i32p]
    f = lib.orc_ivf_search; f.restype = None
    f.argtypes = [_f32p, C.c_void_p, C.c_int, _f32p, C.c_int, _i64p, _i64p, _f32p, C.c_int, C.c_int,
                  C.c_int, C.c_int, C.c_int, _f32p, _i64p, _i32p, C.c_int]
    f = lib.orc_hnsw_create; f.restype = C.c_void_p
    f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int64]
    f = lib.orc_hnsw_free; f.restype = None; f.argtypes = [C.c_void_p]
    f = lib.orc_hnsw_random_level; f.restype = C.c_int; f.argtypes = [C.c_float]
    f = lib.orc_hnsw_build; f.restype = None
    f.argtypes = [C.c_void_p, _f32p, C.c_int64, C.c_void_p, C.c_int]
    f = lib.orc_hnsw_search; f.restype = None
    f.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u32p, _f32p,
                  _i32p, C.c_int]
    f = lib.orc_hnsw_size; f.restype = C.c_int64; f.argtypes = [C.c_void_p]
    f = lib.orc_hnsw_upper_slots; f.restype = C.c_int64; f.argtypes = [C.c_void_p]
    f = lib.orc_hnsw_distance_evals; f.restype = C.c_int64; f.argtypes = []
    f = lib.orc_hnsw_meta; f.restype = None
    f.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    f = lib.orc_hnsw_export; f.restype = None
    f.argtypes = [C.c_void_p, _i32p, _u32p, _i16p, _i64p, _u32p]
    f = lib.orc_recall_at_k; f.restype = C.c_double; f.argtypes = [_i64p, _i64p, C.c_int, C.c_int]
    f = lib.orc_merge_topk; f.restype = None
    f.argtypes = [_f32p, _i64p, C.c_int, C.c_int, C.c_int, _f32p, _i64p]
    f = lib.orc_ml_euclidean; f.restype = C.c_double; f.argtypes = [_f32p, _f32p, C.c_int]
    f = lib.orc_l2_distance_squared; f.restype = C.c_double; f.argtypes = [_f32p, _f32p, C.c_int]
    f = lib.orc_knn_ml; f.restype = None
    f.argtypes = [_f32p, _f64p, C.c_int, C.c_int, _f32p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), _i32p]
    f = lib.orc_kmeanspp_init; f.restype = C.c_int
    f.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _i32p, C.c_int, _i32p]
    f = lib.orc_cluster_minibatch_kmeans; f.restype = C.c_int
    f.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, C.c_int, C.c_int, C.POINTER(C.c_int), _i32p, C.c_void_p]
    f = lib.orc_quantized_row_bytes; f.restype = C.c_int64; f.argtypes = [C.c_int, C.c_int]
    f = lib.orc_quantize_rows; f.restype = None; f.argtypes = [C.c_int, _f32p, C.c_int64, C.c_int, _u8p]
    f = lib.orc_hamming; f.restype = C.c_int; f.argtypes = [_u8p, _u8p, C.c_int]
    f = lib.orc_hamming_knn; f.restype = None
    f.argtypes = [_u8p, C.c_int64, C.c_int, _u8p, C.c_int, C.c_int, _i32p, _i64p]
    f = lib.orc_pq_train; f.restype = C.c_int
    f.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, C.c_int, _f32p]
    f = lib.orc_pq_encode; f.restype = None; f.argtypes = [_f32p, C.c_int64, C.c_int, _f32p, C.c_int, C.c_int, _i16p]
    f = lib.orc_pq_asymmetric_distance; f.restype = C.c_float
    f.argtypes = [_f32p, _i16p, _f32p, C.c_int, C.c_int, C.c_int]
    f = lib.orc_pq_knn; f.restype = None
    f.argtypes = [_f32p, C.c_int, _i16p, C.c_int64, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _i64p, C.c_void_p]
    f = lib.orc_cluster_kmeans; f.restype = C.c_int
    f.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, C.c_int, _i32p, C.c_void_p, C.c_void_p]
    f = lib.orc_fp16_to_float; f.restype = C.c_float; f.argtypes = [C.c_uint16]
    f = lib.orc_keys_from_halfvec; f.restype = None
    f.argtypes = [np.ctypeslib.ndpointer(np.uint16, flags="C"), C.c_int64, C.c_int, _f32p]
    f = lib.orc_keys_from_bits; f.restype = None
    f.argtypes = [np.ctypeslib.ndpointer(np.uint8, flags="C"), C.c_int64, C.c_int, _f32p]
    f = lib.orc_keys_from_sparse; f.restype = None
    f.argtypes = [_i64p, _i32p, _f32p, C.c_int64, C.c_int, _f32p]
    return lib


_LIBS = {}


def lib(native=False):
    name = "libndb_oracle_native.so" if native else "libndb_oracle.so"
    if name not in _LIBS:
        _LIBS[name] = _load(name)
    return _LIBS[name]


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ---- convenience wrappers -------------------------------------------------------------
def distance_pairs(A, B, metric, arith):
    A, B = f32(A), f32(B)
    out = np.empty(A.shape[0], np.float32)
    lib().orc_distance_pairs(A, B, out, A.shape[0], A.shape[1], metric, arith)
    return out


def knn_exact(X, Q, k, metric=L2, arith=ARITH_OP_F64, ids=None, nthreads=8, native=False):
    X, Q = f32(X), f32(Q)
    d = np.empty((Q.shape[0], k), np.float32)
    i = np.empty((Q.shape[0], k), np.int64)
    idp = None if ids is None else np.ascontiguousarray(ids, np.int64).ctypes.data
    lib(native).orc_knn_exact(X, idp, X.shape[0], X.shape[1], Q, Q.shape[0], k, metric, arith, d, i, nthreads)
    return d, i


def kmeans_train(X, k, max_iter=50, threshold=0.001):
    X = f32(X)
    n, dim = X.shape
    Cn = np.zeros((k, dim), np.float32)
    assign = np.zeros(n, np.int32)
    counts = np.zeros(k, np.int32)
    cost = C.c_float(0)
    iters = lib().orc_kmeans_train(X, n, dim, k, max_iter, threshold, Cn, assign, counts, C.byref(cost))
    return Cn, assign, counts, iters, cost.value


def kmeans_assign(X, Cn, nthreads=8):
    X, Cn = f32(X), f32(Cn)
    out = np.empty(X.shape[0], np.int32)
    lib().orc_kmeans_assign(X, X.shape[0], X.shape[1], Cn, Cn.shape[0], out, nthreads)
    return out


def kmeans_update(X, assign, k):
    X = f32(X)
    Cn = np.zeros((k, X.shape[1]), np.float32)
    counts = np.zeros(k, np.int32)
    lib().orc_kmeans_update(X, np.ascontiguousarray(assign, np.int32), X.shape[0], X.shape[1], k, Cn, counts)
    return Cn, counts


def ivf_assign(X, Cn, nthreads=8, native=False):
    X, Cn = f32(X), f32(Cn)
    out = np.empty(X.shape[0], np.int32)
    lib(native).orc_ivf_assign(X, X.shape[0], X.shape[1], Cn, Cn.shape[0], out, nthreads)
    return out


def lists_from_assignment(assign, nlists):
    """CSR over insertion order: what a sequence of ivfinsert calls in row order produces."""
    assign = np.asarray(assign)
    order = np.argsort(assign, kind="stable").astype(np.int64)
    counts = np.bincount(assign, minlength=nlists)
    off = np.zeros(nlists + 1, np.int64)
    off[1:] = np.cumsum(counts)
    return off, order


def ivf_search(X, Cn, list_off, list_rows, Q, nprobe, k, strategy=L2, literal=False, ids=None,
               nthreads=8, native=False):
    X, Cn, Q = f32(X), f32(Cn), f32(Q)
    nq = Q.shape[0]
    d = np.empty((nq, k), np.float32)
    i = np.empty((nq, k), np.int64)
    cnt = np.empty(nq, np.int32)
    idp = None if ids is None else np.ascontiguousarray(ids, np.int64).ctypes.data
    lib(native).orc_ivf_search(X, idp, X.shape[1], Cn, Cn.shape[0],
                               np.ascontiguousarray(list_off, np.int64),
                               np.ascontiguousarray(list_rows, np.int64), Q, nq, nprobe, k,
                               strategy, 1 if literal else 0, d, i, cnt, nthreads)
    return d, i, cnt


def select_clusters(q, Cn, nprobe):
    q, Cn = f32(q), f32(Cn)
    out = np.full(min(nprobe, Cn.shape[0]), -1, np.int32)
    lib().orc_ivf_select_clusters(q, q.shape[0], Cn, Cn.shape[0], nprobe, out)
    return out


class Hnsw:
    def __init__(self, dim, m=16, ef_construction=64, ef_search=40, ml=0.36, capacity=0, native=False):
        self.l = lib(native)
        self.dim, self.m = dim, m
        self.h = self.l.orc_hnsw_create(dim, m, ef_construction, ef_search, ml, capacity)

    def __del__(self):
        if getattr(self, "h", None):
            self.l.orc_hnsw_free(self.h)
            self.h = None

    def build(self, X, levels, mode=0):
        X = f32(X)
        lv = np.ascontiguousarray(levels, np.int32)
        self.l.orc_hnsw_build(self.h, X, X.shape[0], lv.ctypes.data, mode)

    def search(self, Q, ef, k, strategy=L2, search_mode=0, nthreads=8):
        Q = f32(Q)
        nq = Q.shape[0]
        nodes = np.empty((nq, k), np.uint32)
        d = np.empty((nq, k), np.float32)
        cnt = np.empty(nq, np.int32)
        self.l.orc_hnsw_search(self.h, Q, nq, strategy, ef, k, search_mode, nodes, d, cnt, nthreads)
        return d, nodes, cnt

    def distance_evals(self):
        return self.l.orc_hnsw_distance_evals()

    def export(self):
        n = self.l.orc_hnsw_size(self.h)
        levels = np.empty(n, np.int32)
        nbr0 = np.empty((n, 2 * self.m), np.uint32)
        cnt = np.empty((n, 16), np.int16)
        uoff = np.empty(n + 1, np.int64)
        upper = np.empty(max(1, self.l.orc_hnsw_upper_slots(self.h)), np.uint32)
        self.l.orc_hnsw_export(self.h, levels, nbr0, cnt, uoff, upper)
        ep, el, ml = C.c_uint32(), C.c_int(), C.c_int()
        self.l.orc_hnsw_meta(self.h, C.byref(ep), C.byref(el), C.byref(ml))
        return dict(levels=levels, nbr0=nbr0, cnt=cnt, upper_off=uoff, upper=upper,
                    entry_point=ep.value, entry_level=el.value, max_level=ml.value)


def hnsw_levels(n, ml=0.36, seed=42):
    """hnswGetRandomLevel over libc random(), seeded (the reference never seeds: SURVEY Q14)."""
    libc = C.CDLL(None)
    libc.srandom(C.c_uint(seed))
    l = lib()
    return np.array([l.orc_hnsw_random_level(ml) for _ in range(n)], np.int32)


def recall_at_k(found, truth):
    found = np.ascontiguousarray(found, np.int64)
    truth = np.ascontiguousarray(truth, np.int64)
    return lib().orc_recall_at_k(found, truth, found.shape[0], found.shape[1])


def keys_from_halfvec(h):
    h = np.ascontiguousarray(h, np.uint16)
    out = np.empty(h.shape, np.float32)
    lib().orc_keys_from_halfvec(h, h.shape[0], h.shape[1], out)
    return out


def keys_from_bits(bits, nbits):
    bits = np.ascontiguousarray(bits, np.uint8)
    out = np.empty((bits.shape[0], nbits), np.float32)
    lib().orc_keys_from_bits(bits, bits.shape[0], nbits, out)
    return out


def keys_from_sparse(indptr, indices, values, total_dim):
    indptr = np.ascontiguousarray(indptr, np.int64)
    indices = np.ascontiguousarray(indices, np.int32)
    values = f32(values)
    out = np.empty((indptr.shape[0] - 1, total_dim), np.float32)
    lib().orc_keys_from_sparse(indptr, indices, values, indptr.shape[0] - 1, total_dim, out)
    return out


def merge_topk(dist, ids):
    dist = f32(dist); ids = np.ascontiguousarray(ids, np.int64)
    s, nq, k = dist.shape
    od = np.empty((nq, k), np.float32); oi = np.empty((nq, k), np.int64)
    lib().orc_merge_topk(dist, ids, s, nq, k, od, oi)
    return od, oi


# ---- the reference itself (operator distances), compiled under oracle/pgshim ------------
def ref_lib(avx2=False):
    name = "libndb_ref_distance_avx2.so" if avx2 else "libndb_ref_distance.so"
    path = os.path.join(ORACLE_DIR, "_ref", name)
    if not os.path.exists(path):
        return None
    key = "ref:" + name
    if key not in _LIBS:
        l = C.CDLL(path)
        l.ndb_ref_distance.restype = C.c_int
        l.ndb_ref_distance.argtypes = [C.c_int, _f32p, C.c_int, _f32p, C.c_int, C.POINTER(C.c_float)]
        l.ndb_ref_distance_pairs.restype = C.c_int
        l.ndb_ref_distance_pairs.argtypes = [C.c_int, _f32p, _f32p, _f32p, C.c_long, C.c_int]
        l.ndb_ref_last_error.restype = C.c_char_p
        l.ndb_ref_table_create.restype = C.c_void_p
        l.ndb_ref_table_create.argtypes = [_f32p, C.c_long, C.c_int]
        l.ndb_ref_table_free.restype = None
        l.ndb_ref_table_free.argtypes = [C.c_void_p]
        l.ndb_ref_seqscan.restype = C.c_int
        l.ndb_ref_seqscan.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p]
        _LIBS[key] = l
    return _LIBS[key]


def ref_distance(metric, a, b, avx2=False):
    """Returns (rc, value, error message) from the reference's fmgr function."""
    l = ref_lib(avx2)
    a, b = f32(a), f32(b)
    out = C.c_float(0)
    rc = l.ndb_ref_distance(metric, a if a.size else np.zeros(1, np.float32), a.size,
                            b if b.size else np.zeros(1, np.float32), b.size, C.byref(out))
    return rc, out.value, (l.ndb_ref_last_error() or b"").decode()


def ref_distance_pairs(metric, A, B, avx2=False):
    l = ref_lib(avx2)
    A, B = f32(A), f32(B)
    out = np.empty(A.shape[0], np.float32)
    rc = l.ndb_ref_distance_pairs(metric, A, B, out, A.shape[0], A.shape[1])
    assert rc == 0, l.ndb_ref_last_error()
    return out


# ---- relation images in the reference's page layout (oracle/ndb_oracle_pages.c) ------------
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def _pages_lib():
    l = lib()
    if not getattr(l, "_pages_bound", False):
        l.orc_ivf_encode_relation.restype = C.c_int64
        l.orc_ivf_encode_relation.argtypes = [_f32p, C.c_void_p, C.c_int64, C.c_int, _f32p, C.c_int, C.c_int, _i32p, _u8p,
                                              C.c_int64, C.c_int]
        l.orc_page_mark_dead.restype = None
        l.orc_page_mark_dead.argtypes = [_u8p, C.c_int64, C.c_int]
        l.orc_hnsw_encode_relation.restype = C.c_int64
        l.orc_hnsw_encode_relation.argtypes = [C.c_void_p, _f32p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p,
                                               C.c_int64]
        l._pages_bound = True
    return l


def ivf_encode_relation(X, Cn, assign, tids=None, nprobe=10, multi_page_centroids=True):
    """Blocks (uint8 [nblocks, 8192]) as ivfbuild + ivfinsert-per-row would leave them."""
    X, Cn = f32(X), f32(Cn)
    n, dim = X.shape
    per_page = max(1, (8192 - 24 - 8) // (8 + 4 * dim + 4))
    cap = 2 + Cn.shape[0] + n // per_page + 2 * Cn.shape[0] + 8
    blocks = np.zeros((cap, 8192), np.uint8)
    tp = None if tids is None else np.ascontiguousarray(tids, np.int64).ctypes.data
    nb = _pages_lib().orc_ivf_encode_relation(X, tp, n, dim, Cn, Cn.shape[0], nprobe,
                                              np.ascontiguousarray(assign, np.int32), blocks.reshape(-1), cap,
                                              1 if multi_page_centroids else 0)
    if nb < 0:
        return nb, None
    return nb, blocks[:nb]


def page_mark_dead(blocks, block, offnum):
    _pages_lib().orc_page_mark_dead(blocks.reshape(-1), block, offnum)


def hnsw_encode_relation(g, X, tids=None, efc=64, efs=40):
    X = f32(X)
    n = X.shape[0]
    blocks = np.zeros((n + 1, 8192), np.uint8)
    tp = None if tids is None else np.ascontiguousarray(tids, np.int64).ctypes.data
    nb = _pages_lib().orc_hnsw_encode_relation(g.h, X, tp, g.dim, g.m, efc, efs, blocks.reshape(-1), n + 1)
    return nb, blocks


def ref_fp16_lib():
    """fp16_to_float cut out of the reference's quantization.c (oracle/Makefile ref_build), or None."""
    path = os.path.join(ORACLE_DIR, "_ref", "libndb_ref_fp16.so")
    if not os.path.exists(path):
        return None
    lib_ = C.CDLL(path)
    lib_.fp16_to_float.restype = C.c_float
    lib_.fp16_to_float.argtypes = [C.c_uint16]
    return lib_


def ref_leafs_lib():
    """The reference's own ivfComputeDistance, hnswComputeDistance, hnswGetRandomLevel and k-means block
    (kmeans_init ... find_nearest_centroid), cut out of ivf_am.c / hnsw_am.c by oracle/extract_ref_leafs.py
    and compiled with the reference's flags; None when oracle/_ref has not been built."""
    path = os.path.join(ORACLE_DIR, "_ref", "libndb_ref_leafs.so")
    if not os.path.exists(path):
        return None
    l = C.CDLL(path)
    l.ref_ivf_distance.restype = C.c_float; l.ref_ivf_distance.argtypes = [_f32p, _f32p, C.c_int, C.c_int]
    l.ref_hnsw_distance.restype = C.c_float; l.ref_hnsw_distance.argtypes = [_f32p, _f32p, C.c_int, C.c_int]
    l.ref_hnsw_random_level.restype = C.c_int; l.ref_hnsw_random_level.argtypes = [C.c_float]
    l.ref_kmeans_train.restype = None
    l.ref_kmeans_train.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _f32p, _i32p, _i32p]
    l.ref_layout.restype = C.c_int
    l.ref_layout.argtypes = [_i64p]
    l.ref_knn_ml.restype = None
    l.ref_knn_ml.argtypes = [_f32p, _f64p, C.c_int, C.c_int, _f32p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_void_p]
    l.ref_ml_euclidean.restype = C.c_double; l.ref_ml_euclidean.argtypes = [_f32p, _f32p, C.c_int]
    l.ref_l2_distance_squared.restype = C.c_double; l.ref_l2_distance_squared.argtypes = [_f32p, _f32p, C.c_int]
    l.ref_kmeanspp_init.restype = None; l.ref_kmeanspp_init.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _i32p]
    l.ref_cluster_kmeans.restype = C.c_int
    l.ref_cluster_kmeans.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, C.c_void_p]
    l.ref_cluster_minibatch_kmeans.restype = None
    l.ref_cluster_minibatch_kmeans.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, C.c_void_p]
    l.ref_quantize_row.restype = C.c_int; l.ref_quantize_row.argtypes = [C.c_int, _f32p, C.c_int, _u8p, C.c_int]
    l.ref_hamming.restype = C.c_int; l.ref_hamming.argtypes = [_u8p, _u8p, C.c_int]
    l.ref_pq_train_subspace.restype = None
    l.ref_pq_train_subspace.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int]
    l.ref_pq_encode.restype = None; l.ref_pq_encode.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, _i16p]
    l.ref_pq_asymmetric_distance.restype = C.c_float
    l.ref_pq_asymmetric_distance.argtypes = [_f32p, _i16p, _f32p, C.c_int, C.c_int, C.c_int]
    return l


def ref_page_layout():
    """Sizes / field offsets of the reference's own on-page structs (IvfMetaPageData ... HnswNodeData)."""
    o = np.zeros(64, np.int64)
    n = ref_leafs_lib().ref_layout(o)
    return o[:n].copy()


def page_layout():
    """The same list as the oracle's relation encoders use it (ndb_oracle_pages.c)."""
    o = np.zeros(64, np.int64)
    f = lib().orc_page_layout
    f.restype = C.c_int
    f.argtypes = [_i64p]
    n = f(o)
    return o[:n].copy()


def ref_kmeans_train(X, k):
    X = f32(X)
    n, d = X.shape
    Cn = np.zeros((k, d), np.float32)
    assign = np.zeros(n, np.int32)
    counts = np.zeros(k, np.int32)
    ref_leafs_lib().ref_kmeans_train(X, n, d, k, Cn, assign, counts)
    return Cn, assign, counts


# ---- knn_classify / knn_regress / cluster_kmeans (oracle/ndb_oracle_ml.c) --------------------------------------
def libc_rand_draws(seed, k):
    """What rand() returns k times after srand(seed) in this libc (cluster_kmeans' kmeanspp_init draws exactly k)."""
    libc = C.CDLL(None)
    libc.srand(C.c_uint(seed))
    libc.rand.restype = C.c_int
    return np.array([libc.rand() for _ in range(k)], np.int32)


def knn_ml(X, labels, Q, k):
    """(class, mean, rows[k]) per query: the oracle's knn_classify / knn_regress over the resident rows."""
    X, Q = f32(X), f32(Q)
    labels = np.ascontiguousarray(labels, np.float64)
    n, d = X.shape
    cls, mean, rows = np.zeros(len(Q), np.int32), np.zeros(len(Q)), np.zeros((len(Q), min(k, n)), np.int32)
    for j in range(len(Q)):
        c, m = C.c_int(), C.c_double()
        lib().orc_knn_ml(X, labels, n, d, Q[j], k, C.byref(c), C.byref(m), rows[j])
        cls[j], mean[j] = c.value, m.value
    return cls, mean, rows


def ref_knn_ml(X, labels, Q, k):
    """The same through the reference's own KNNSample / euclidean_distance / compare_samples / qsort; also the
    n distances per query."""
    X, Q = f32(X), f32(Q)
    labels = np.ascontiguousarray(labels, np.float64)
    n, d = X.shape
    cls, mean, dist = np.zeros(len(Q), np.int32), np.zeros(len(Q)), np.zeros((len(Q), n))
    for j in range(len(Q)):
        c, m = C.c_int(), C.c_double()
        ref_leafs_lib().ref_knn_ml(X, labels, n, d, Q[j], k, C.byref(c), C.byref(m), dist[j].ctypes.data)
        cls[j], mean[j] = c.value, m.value
    return cls, mean, dist


def cluster_kmeans(X, k, max_iters, draws):
    """(labels 1-based, centers, seeds, iterations): oracle restatement of cluster_kmeans with the given rand() draws."""
    X = f32(X)
    n, d = X.shape
    labels, centers, seeds = np.zeros(n, np.int32), np.zeros((k, d), np.float32), np.zeros(k, np.int32)
    draws = np.ascontiguousarray(draws, np.int32)
    it = lib().orc_cluster_kmeans(X, n, d, k, max_iters, draws, RAND_MAX, labels, centers.ctypes.data, seeds.ctypes.data)
    return labels, centers, seeds, it


def ref_cluster_kmeans(X, k, max_iters, seed):
    """The reference's kmeanspp_init + Lloyd loop after srand(seed): (labels, centers, iterations)."""
    X = f32(X)
    n, d = X.shape
    labels, centers = np.zeros(n, np.int32), np.zeros((k, d), np.float32)
    C.CDLL(None).srand(C.c_uint(seed))
    it = ref_leafs_lib().ref_cluster_kmeans(X, n, d, k, max_iters, labels, centers.ctypes.data)
    return labels, centers, it


# ---- product quantisation (oracle/ndb_oracle_ml.c; ml_product_quantization.c) ------------------------------------
def pq_train(X, m, ksub, draws, max_iters=100):
    """train_pq_codebook's per-subspace k-means: codebooks [m][ksub][dsub]; draws = the m*ksub values of rand()."""
    X = f32(X)
    n, dim = X.shape
    cb = np.zeros((m, ksub, dim // m), np.float32)
    rc = lib().orc_pq_train(X, n, dim, m, ksub, np.ascontiguousarray(draws, np.int32), max_iters, cb.reshape(-1))
    assert rc == 0
    return cb


def pq_encode(X, cb):
    X = f32(X)
    m, ksub, dsub = cb.shape
    codes = np.zeros((len(X), m), np.int16)
    lib().orc_pq_encode(X, len(X), X.shape[1], f32(cb).reshape(-1), m, ksub, codes.reshape(-1))
    return codes


def pq_knn(Q, codes, cb, k, want_all=False):
    """(dist [nq,k], rows [nq,k], all distances or None): ORDER BY pq_asymmetric_distance LIMIT k, ties by row."""
    Q = f32(Q)
    m, ksub, dsub = cb.shape
    codes = np.ascontiguousarray(codes, np.int16)
    n = len(codes)
    d, r = np.zeros((len(Q), k), np.float32), np.zeros((len(Q), k), np.int64)
    al = np.zeros((len(Q), n), np.float32) if want_all else None
    lib().orc_pq_knn(Q, len(Q), codes.reshape(-1), n, f32(cb).reshape(-1), m * dsub, m, ksub, k, d.reshape(-1), r.reshape(-1),
                     None if al is None else al.ctypes.data)
    return d, r, al


def ref_pq_train(X, m, ksub, seed, max_iters=100):
    """The reference's own train_subspace_kmeans per subspace after srand(seed), in train_pq_codebook's order."""
    X = f32(X)
    n, dim = X.shape
    dsub = dim // m
    cb = np.zeros((m, ksub, dsub), np.float32)
    C.CDLL(None).srand(C.c_uint(seed))
    for sub in range(m):
        S = np.ascontiguousarray(X[:, sub * dsub:(sub + 1) * dsub])
        ref_leafs_lib().ref_pq_train_subspace(S.reshape(-1), n, dsub, ksub, cb[sub].reshape(-1), max_iters)
    return cb


def ref_pq_encode(X, cb):
    X = f32(X)
    m, ksub, dsub = cb.shape
    codes = np.zeros((len(X), m), np.int16)
    flat = f32(cb).reshape(-1)
    for i in range(len(X)):
        ref_leafs_lib().ref_pq_encode(X[i], flat, m, ksub, dsub, codes[i])
    return codes


def ref_pq_distances(Q, codes, cb):
    Q = f32(Q)
    m, ksub, dsub = cb.shape
    codes = np.ascontiguousarray(codes, np.int16)
    flat = f32(cb).reshape(-1)
    out = np.zeros((len(Q), len(codes)), np.float32)
    f = ref_leafs_lib().ref_pq_asymmetric_distance
    for j in range(len(Q)):
        for i in range(len(codes)):
            out[j, i] = f(Q[j], codes[i], flat, m, ksub, dsub)
    return out


# ---- per-vector quantisers and the Hamming scan (oracle/ndb_oracle_ml.c; src/types/quantization.c) ----------------
def quantize_rows(kind, X):
    X = f32(X)
    n, dim = X.shape
    rb = lib().orc_quantized_row_bytes(kind, dim)
    out = np.zeros((n, rb), np.uint8)
    lib().orc_quantize_rows(kind, X, n, dim, out.reshape(-1))
    return out


def ref_quantize_rows(kind, X):
    """The reference's own quantize_vector_* per row (data[] bytes)."""
    X = f32(X)
    n, dim = X.shape
    rb = lib().orc_quantized_row_bytes(kind, dim)
    out = np.zeros((n, rb), np.uint8)
    for i in range(n):
        assert ref_leafs_lib().ref_quantize_row(kind, X[i], dim, out[i], rb) == 0
    return out


def hamming_knn(rows, nbits, Q, k):
    rows, Q = np.ascontiguousarray(rows, np.uint8), np.ascontiguousarray(Q, np.uint8)
    d, i = np.zeros((len(Q), k), np.int32), np.zeros((len(Q), k), np.int64)
    lib().orc_hamming_knn(rows.reshape(-1), len(rows), nbits, Q.reshape(-1), len(Q), k, d.reshape(-1), i.reshape(-1))
    return d, i


# ---- cluster_minibatch_kmeans (oracle/ndb_oracle_ml.c; ml_minibatch_kmeans.c) --------------------------------------
def cluster_minibatch_kmeans(X, k, batch_size, max_iters, draws):
    """(labels 1-based, centers, draws consumed): oracle restatement with the given rand() values."""
    X = f32(X)
    n, d = X.shape
    labels, centers = np.zeros(n, np.int32), np.zeros((k, d), np.float32)
    draws = np.ascontiguousarray(draws, np.int32)
    used = C.c_int()
    rc = lib().orc_cluster_minibatch_kmeans(X, n, d, k, batch_size, max_iters, draws, len(draws), RAND_MAX, C.byref(used), labels,
                                            centers.ctypes.data)
    assert rc == 0, rc
    return labels, centers, used.value


def ref_cluster_minibatch_kmeans(X, k, batch_size, max_iters, seed):
    """The reference's minibatch_kmeans_pp_init + main loop after srand(seed): (labels, centers)."""
    X = f32(X)
    n, d = X.shape
    labels, centers = np.zeros(n, np.int32), np.zeros((k, d), np.float32)
    C.CDLL(None).srand(C.c_uint(seed))
    ref_leafs_lib().ref_cluster_minibatch_kmeans(X, n, d, k, batch_size, max_iters, labels, centers.ctypes.data)
    return labels, centers
