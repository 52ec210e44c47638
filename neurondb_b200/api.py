"""Host-side mirror of the reference's interface for the vector-search hot path.

Names follow the reference (NeuronDB/src/index/ivf_am.c, hnsw_am.c, opclass.c): the operator
procedures, `ivfbuild` / `ivfinsert` / `ivfgettuple`-style index objects.  Everything here calls
the C ABI in include/ndb_b200.h; nothing is computed in Python and there is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from ._lib import (ARITH_AVX2, ARITH_AVX512, ARITH_FAST, ARITH_HNSW, ARITH_IVF_F32, ARITH_OP_F64, ARITH_TENSOR, COSINE,
                   HNSW_BESTFIRST, HNSW_LITERAL,
                   HNSW_SELECT_CLOSEST, HNSW_SELECT_HEURISTIC, IP,
                   IVF_FULL, IVF_LITERAL, L2, NdbError, check, f32, ptr,
                   QUANT_BINARY, QUANT_FP16, QUANT_INT4, QUANT_INT8, QUANT_TERNARY, QUANT_UINT8)

_initialised = {"device": None}


def init(device=0):
    check(L.load().ndb_b200_init(device))
    _initialised["device"] = device


def shutdown():
    L.load().ndb_b200_shutdown()
    _initialised["device"] = None


def is_available():
    return bool(L.load().ndb_b200_is_available())


def launch_count():
    return int(L.load().ndb_b200_launch_count())


def set_timing(on):
    check(L.load().ndb_b200_set_timing(1 if on else 0))


def last_kernel_stats():
    ms, b, ev = C.c_double(), C.c_double(), C.c_int64()
    check(L.load().ndb_b200_last_kernel_stats(C.byref(ms), C.byref(b), C.byref(ev)))
    return ms.value, b.value, ev.value


# ---- operator procedures (opclass.c:53-161): <->, <=>, <#> over n pairs ------------------------
def distance_pairs(A, B, metric=L2, arith=ARITH_OP_F64):
    A, B = f32(A), f32(B)
    if A.ndim == 1:
        A, B = A[None, :], B[None, :]
    if A.shape != B.shape:
        raise NdbError(-5, "vector dimensions must match: %s vs %s" % (A.shape, B.shape))
    out = np.empty(A.shape[0], np.float32)
    check(L.load().ndb_b200_distance_pairs(metric, arith, ptr(A), ptr(B), ptr(out), A.shape[0], A.shape[1]))
    return out


def vector_l2_distance_op(a, b):
    return distance_pairs(a, b, L2)


def vector_cosine_distance_op(a, b):
    return distance_pairs(a, b, COSINE)


def vector_inner_product_distance_op(a, b):
    return distance_pairs(a, b, IP)


def distance_rows(X, q, metric=L2, arith=ARITH_OP_F64):
    X, q = f32(X), f32(q)
    out = np.empty(X.shape[0], np.float32)
    check(L.load().ndb_b200_distance_rows(metric, arith, ptr(X), X.shape[0], X.shape[1], ptr(q), ptr(out)))
    return out


def vector_distance_batch(vectors, query, metric=L2):
    """vector_{l2,cosine,inner_product}_distance_batch(vector[], vector) (vector_batch.c:37-420): `vectors` is a list
    whose elements are 1-d arrays or None; returns (distances float32[n], nulls bool[n])."""
    q = f32(query).reshape(-1)
    dim = q.shape[0]
    n = len(vectors)
    rows = np.zeros((max(n, 1), dim), np.float32)
    dims = np.zeros(max(n, 1), np.int32)
    for i, v in enumerate(vectors):
        if v is None:
            continue
        v = f32(v).reshape(-1)
        dims[i] = v.shape[0]
        if v.shape[0] == dim:
            rows[i] = v
    out = np.zeros(max(n, 1), np.float32)
    nulls = np.zeros(max(n, 1), np.uint8)
    check(L.load().ndb_b200_vector_distance_batch(metric, ptr(rows), ptr(dims), n, dim, ptr(q), ptr(out), ptr(nulls)))
    return out[:n], nulls[:n].astype(bool)


# ---- ndb_gpu_backend launchers (neurondb_gpu_backend.h:54-79) -----------------------------------
def launch_l2_distance(A, B):
    A, B = f32(A), f32(B)
    out = np.empty(A.shape[0], np.float32)
    check(L.load().ndb_b200_launch_l2_distance(ptr(A), ptr(B), ptr(out), A.shape[0], A.shape[1], None))
    return out


def launch_cosine(A, B):
    A, B = f32(A), f32(B)
    out = np.empty(A.shape[0], np.float32)
    check(L.load().ndb_b200_launch_cosine(ptr(A), ptr(B), ptr(out), A.shape[0], A.shape[1], None))
    return out


def launch_kmeans_assign(X, Cn):
    X, Cn = f32(X), f32(Cn)
    idx = np.empty(X.shape[0], np.int32)
    check(L.load().ndb_b200_launch_kmeans_assign(ptr(X), ptr(Cn), ptr(idx), X.shape[0], X.shape[1], Cn.shape[0], None))
    return idx


def launch_kmeans_update(X, idx, k):
    X = f32(X)
    idx = np.ascontiguousarray(idx, np.int32)
    Cn = np.zeros((k, X.shape[1]), np.float32)
    check(L.load().ndb_b200_launch_kmeans_update(ptr(X), ptr(idx), ptr(Cn), X.shape[0], X.shape[1], k, None))
    return Cn


def kmeans_train(X, k, max_iter=50, tol=0.001):
    """kmeans_init + kmeans_run (ivf_am.c:2070-2159). Returns (C, assign, counts, iters, cost)."""
    X = f32(X)
    n, d = X.shape
    Cn = np.zeros((k, d), np.float32)
    assign = np.zeros(n, np.int32)
    counts = np.zeros(k, np.int32)
    iters, cost = C.c_int(), C.c_float()
    check(L.load().ndb_b200_kmeans_train(ptr(X), n, d, k, max_iter, tol, ptr(Cn), ptr(assign), ptr(counts),
                                         C.byref(iters), C.byref(cost)))
    return Cn, assign, counts, iters.value, cost.value


def cluster_kmeans(X, k, max_iters, rand_draws, rand_max=2147483647):
    """cluster_kmeans (ml_kmeans.c:146-303): returns (labels 1-based, centers, seeds, iterations).  rand_draws = the k
    values rand() returned, in call order."""
    X = f32(X)
    if X.ndim != 2:
        raise L.NdbError(-1, "cluster_kmeans: a 2-d array of rows expected")
    n, d = X.shape
    draws = np.ascontiguousarray(rand_draws, np.int32)
    if draws.shape != (max(k, 0),):
        raise L.NdbError(-1, "cluster_kmeans: one rand() value per cluster expected")
    labels = np.zeros(n, np.int32)
    Cn = np.zeros((max(k, 0), d), np.float32)
    seeds = np.zeros(max(k, 0), np.int32)
    iters = C.c_int()
    check(L.load().ndb_b200_cluster_kmeans(ptr(X), n, d, k, max_iters, ptr(draws), rand_max, ptr(labels), ptr(Cn),
                                           C.byref(iters), ptr(seeds)))
    return labels, Cn, seeds, iters.value


def merge_topk(dist, ids):
    dist = f32(dist)
    ids = np.ascontiguousarray(ids, np.int64)
    s, nq, k = dist.shape
    od = np.empty((nq, k), np.float32)
    oi = np.empty((nq, k), np.int64)
    check(L.load().ndb_b200_merge_topk(ptr(dist), ptr(ids), s, nq, k, ptr(od), ptr(oi)))
    return od, oi


def keys_from_vector(datums, dim):
    """Detoasted `vector` datums (bytes: n x (8-byte varlena header + 4*dim)) -> float32 rows."""
    buf = np.frombuffer(datums, np.uint8) if not isinstance(datums, np.ndarray) else np.ascontiguousarray(datums, np.uint8).reshape(-1)
    stride = 8 + 4 * dim
    if buf.size % stride:
        raise ValueError("datums must be a whole number of %d-byte Vector values" % stride)
    n = buf.size // stride
    out = np.empty((n, dim), np.float32)
    check(L.load().ndb_b200_keys_from_vector(ptr(buf), n, dim, ptr(out)))
    return out


def keys_from_halfvec(h):
    """halfvec keys (uint16 IEEE binary16 bit patterns, [n, dim]) -> float32 rows (hnsw_am.c:1435-1450)."""
    h = np.ascontiguousarray(h, np.uint16)
    n, dim = h.shape
    out = np.empty((n, dim), np.float32)
    check(L.load().ndb_b200_keys_from_halfvec(ptr(h), n, dim, ptr(out)))
    return out


def keys_from_bits(bits, nbits):
    """bit keys ([n, ceil(nbits/8)] bytes, MSB first) -> rows of +1/-1 (hnsw_am.c:1480-1507)."""
    bits = np.ascontiguousarray(bits, np.uint8)
    n = bits.shape[0]
    if bits.shape[1] != (nbits + 7) // 8:
        raise ValueError("bits must have ceil(nbits/8) bytes per row")
    out = np.empty((n, nbits), np.float32)
    check(L.load().ndb_b200_keys_from_bits(ptr(bits), n, nbits, ptr(out)))
    return out


def keys_from_sparse(indptr, indices, values, total_dim):
    """sparsevec keys as a CSR batch -> dense float32 rows (hnsw_am.c:1451-1479)."""
    indptr = np.ascontiguousarray(indptr, np.int64)
    indices = np.ascontiguousarray(indices, np.int32)
    values = f32(values)
    n = indptr.shape[0] - 1
    out = np.empty((n, total_dim), np.float32)
    check(L.load().ndb_b200_keys_from_sparse(ptr(indptr), ptr(indices), ptr(values), n, total_dim, ptr(out)))
    return out


def _rows(a, dim, what):
    """float32 C-contiguous [n, dim] view of `a`; a wrong shape is the reference's dimension error (EDIM)."""
    a = f32(a)
    if a.ndim != 2 or a.shape[1] != dim:
        raise NdbError(-5, "%s: expected [n, %d] float32 rows, got shape %s" % (what, dim, a.shape))
    return a


def _pinned_like(a, dtype, shape, what):
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags.c_contiguous or tuple(a.shape) != tuple(shape):
        raise NdbError(-1, "%s: expected a C-contiguous %s array of shape %s" % (what, np.dtype(dtype).name, tuple(shape)))
    return a


# ---- multi-GPU communicator (ndb_b200_comm_*) -----------------------------------------------------
def comm_unique_id():
    buf = (C.c_ubyte * 128)()
    check(L.load().ndb_b200_comm_unique_id(buf, 128))
    return bytes(buf)


def comm_init(rank, world, uid):
    buf = (C.c_ubyte * 128).from_buffer_copy(uid) if world > 1 else None
    check(L.load().ndb_b200_comm_init(rank, world, buf, 128))


def comm_init_torch():
    """One process per GPU under torchrun: rank 0 draws the NCCL id, torch.distributed (any backend) hands it
    to the other ranks, every rank joins.  torch is plumbing here; the data path is the library's own
    communicator."""
    import torch
    import torch.distributed as dist
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    uid = [comm_unique_id() if rank == 0 and world > 1 else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=0)
    comm_init(rank, world, uid[0])
    return rank, world


def comm_shutdown():
    check(L.load().ndb_b200_comm_shutdown())


def comm_nranks():
    return int(L.load().ndb_b200_comm_nranks())


def comm_rank():
    return int(L.load().ndb_b200_comm_rank())


def kmeans_train_sharded_dev(x_ptr, n_local, d, k, c_ptr, assign_ptr, counts_ptr=None, max_iter=50, tol=0.001, stream=None):
    """ndb_b200_kmeans_train_sharded_dev on device pointers; returns (iters, cost)."""
    iters, cost = C.c_int(), C.c_float()
    check(L.load().ndb_b200_kmeans_train_sharded_dev(ptr(x_ptr), n_local, d, k, max_iter, tol, ptr(c_ptr), ptr(assign_ptr),
                                                     ptr(counts_ptr), C.byref(iters), C.byref(cost), ptr(stream)))
    return iters.value, cost.value


class _Handle:
    _free = None

    def __init__(self):
        self.h = C.c_void_p()

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            getattr(L.load(), self._free)(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


RAND_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)          # ndb_b200_rand_fn


def cluster_minibatch_kmeans(X, k, batch_size, max_iters, rand_draws, rand_max=2147483647):
    """cluster_minibatch_kmeans (ml_minibatch_kmeans.c:206-449): (labels 1-based, centers, rand() values consumed).  The
    library calls back for every rand() the reference would make; here the values come from `rand_draws` in order."""
    X = f32(X)
    if X.ndim != 2:
        raise L.NdbError(-1, "cluster_minibatch_kmeans: a 2-d array of rows expected")
    n, d = X.shape
    draws = [int(v) for v in np.asarray(rand_draws).reshape(-1)]
    used = [0]

    def nxt(_state):
        i = used[0]
        used[0] += 1
        return draws[i] if i < len(draws) else -1

    cb = RAND_FN(nxt)
    labels = np.zeros(n, np.int32)
    Cn = np.zeros((max(k, 0), d), np.float32)
    check(L.load().ndb_b200_cluster_minibatch_kmeans(ptr(X), n, d, k, batch_size, max_iters, cb, None, rand_max, ptr(labels), ptr(Cn)))
    return labels, Cn, used[0]


def quantize_rows(kind, X):
    """quantize_vector_i8 / _f16 / _binary / _uint8 / _ternary / _int4 (src/types/quantization.c) per row: uint8 [n][row_bytes]."""
    X = f32(X)
    if X.ndim != 2:
        raise L.NdbError(-1, "quantize_rows: a 2-d array of rows expected")
    n, dim = X.shape
    rb = int(L.load().ndb_b200_quantized_row_bytes(kind, dim))
    if rb <= 0:
        raise L.NdbError(-1, "quantize_rows: unknown kind %d" % kind)
    out = np.zeros((n, rb), np.uint8)
    check(L.load().ndb_b200_quantize_rows(kind, ptr(X), n, dim, ptr(out)))
    return out


def hamming_knn(rows, nbits, Q, k):
    """ORDER BY binary_hamming_distance(bits, q) LIMIT k: (dist int32 [nq][k], rows int64 [nq][k])."""
    rows, Q = np.ascontiguousarray(rows, np.uint8), np.ascontiguousarray(Q, np.uint8)
    nbytes = (nbits + 7) // 8
    if rows.ndim != 2 or Q.ndim != 2 or rows.shape[1] != nbytes or Q.shape[1] != nbytes:
        raise L.NdbError(-5, "binary vector dimensions must match")
    d = np.empty((Q.shape[0], k), np.int32)
    i = np.empty((Q.shape[0], k), np.int64)
    check(L.load().ndb_b200_hamming_knn(ptr(rows), rows.shape[0], nbits, ptr(Q), Q.shape[0], k, ptr(d), ptr(i)))
    return d, i


def _pq_shape(cb):
    cb = f32(cb)
    if cb.ndim != 3:
        raise L.NdbError(-1, "pq: codebooks are [m][ksub][dsub]")
    m, ksub, dsub = cb.shape
    return cb, m, ksub, dsub


def pq_train(X, m, ksub, rand_draws, max_iters=100):
    """train_pq_codebook (ml_product_quantization.c:195-415): codebooks [m][ksub][dim/m]; rand_draws = m*ksub rand() values."""
    X = f32(X)
    if X.ndim != 2:
        raise L.NdbError(-1, "pq_train: a 2-d array of rows expected")
    n, dim = X.shape
    draws = np.ascontiguousarray(rand_draws, np.int32)
    if draws.size != max(m, 0) * max(ksub, 0):
        raise L.NdbError(-1, "pq_train: m * ksub rand() values expected")
    cb = np.zeros((max(m, 1), max(ksub, 1), max(1, dim // max(m, 1))), np.float32)
    check(L.load().ndb_b200_pq_train(ptr(X), n, dim, m, ksub, max_iters, ptr(draws), ptr(cb)))
    return cb


def pq_encode(X, codebooks):
    """pq_encode_vector (:421-536) for every row: int16 codes [n][m]."""
    cb, m, ksub, dsub = _pq_shape(codebooks)
    X = _rows(X, m * dsub, "pq_encode")
    codes = np.zeros((X.shape[0], m), np.int16)
    check(L.load().ndb_b200_pq_encode(ptr(X), X.shape[0], m * dsub, ptr(cb), m, ksub, ptr(codes)))
    return codes


def launch_pq_encode(X, codebooks):
    """The backend vtable's launch_pq_encode: byte codes [n][m]."""
    cb, m, ksub, dsub = _pq_shape(codebooks)
    X = _rows(X, m * dsub, "launch_pq_encode")
    codes = np.zeros((X.shape[0], m), np.uint8)
    check(L.load().ndb_b200_launch_pq_encode(ptr(X), ptr(cb), ptr(codes), X.shape[0], m * dsub, m, ksub, None))
    return codes


class PqIndex(_Handle):
    """Encoded rows resident on the device + ORDER BY pq_asymmetric_distance(q, codes, codebook) LIMIT k (:1003-1110)."""
    _free = "ndb_b200_pq_free"

    def __init__(self, codebooks):
        super().__init__()
        cb, self.m, self.ksub, self.dsub = _pq_shape(codebooks)
        self.dim = self.m * self.dsub
        check(L.load().ndb_b200_pq_create(self.dim, self.m, self.ksub, ptr(cb), C.byref(self.h)))

    def __len__(self):
        return int(L.load().ndb_b200_pq_size(self.h))

    def add(self, X, want_codes=False):
        X = _rows(X, self.dim, "pq_add")
        codes = np.zeros((X.shape[0], self.m), np.int16) if want_codes else None
        check(L.load().ndb_b200_pq_add(self.h, ptr(X), X.shape[0], ptr(codes)))
        return codes

    def add_codes(self, codes):
        codes = np.ascontiguousarray(codes, np.int16)
        if codes.ndim != 2 or codes.shape[1] != self.m:
            raise L.NdbError(-5, "pq_add_codes: codes are [n][m]")
        check(L.load().ndb_b200_pq_add_codes(self.h, ptr(codes), codes.shape[0]))

    def search(self, Q, k):
        Q = _rows(Q, self.dim, "pq_search")
        d = np.empty((Q.shape[0], k), np.float32)
        r = np.empty((Q.shape[0], k), np.int64)
        check(L.load().ndb_b200_pq_search(self.h, ptr(Q), Q.shape[0], k, ptr(d), ptr(r)))
        return d, r

    def distances(self, Q):
        Q = _rows(Q, self.dim, "pq_distances")
        d = np.empty((Q.shape[0], len(self)), np.float32)
        cnt = C.c_ulonglong()
        check(L.load().ndb_b200_pq_distances(self.h, ptr(Q), Q.shape[0], ptr(d), C.byref(cnt)))
        return d, cnt.value


class Dataset(_Handle):
    """The heap column a SeqScan reads; exact kNN = ORDER BY v <op> q LIMIT k."""
    _free = "ndb_b200_dataset_free"

    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        check(L.load().ndb_b200_dataset_create(dim, C.byref(self.h)))

    def append(self, rows, ids=None):
        rows = _rows(rows, self.dim, "dataset_append")
        idp = None if ids is None else np.ascontiguousarray(ids, np.int64)
        check(L.load().ndb_b200_dataset_append(self.h, ptr(rows), ptr(idp), rows.shape[0]))

    def append_dev(self, rows_ptr, n, ids_ptr=None, stream=None):
        check(L.load().ndb_b200_dataset_append_dev(self.h, ptr(rows_ptr), ptr(ids_ptr), n, ptr(stream)))

    def __len__(self):
        return int(L.load().ndb_b200_dataset_size(self.h))

    def knn(self, Q, k, metric=L2, arith=ARITH_OP_F64):
        Q = _rows(Q, self.dim, "knn_exact")
        d = np.empty((Q.shape[0], k), np.float32)
        i = np.empty((Q.shape[0], k), np.int64)
        check(L.load().ndb_b200_knn_exact(self.h, metric, arith, ptr(Q), Q.shape[0], k, ptr(d), ptr(i)))
        return d, i

    def knn_classify(self, labels, Q, k):
        """knn_classify (ml_knn.c:112-357) for a batch of query vectors: class 0 / 1 per query."""
        Q = _rows(Q, self.dim, "knn_classify")
        lab = np.ascontiguousarray(labels, np.float64)
        if lab.shape != (len(self),):
            raise NdbError(-1, "knn_classify: one label per row expected")
        out = np.empty(Q.shape[0], np.int32)
        check(L.load().ndb_b200_knn_classify(self.h, ptr(lab), ptr(Q), Q.shape[0], k, ptr(out)))
        return out

    def knn_regress(self, targets, Q, k):
        """knn_regress (ml_knn.c:363-569): mean target of the k nearest rows per query."""
        Q = _rows(Q, self.dim, "knn_regress")
        t = np.ascontiguousarray(targets, np.float64)
        if t.shape != (len(self),):
            raise NdbError(-1, "knn_regress: one target per row expected")
        out = np.empty(Q.shape[0], np.float64)
        check(L.load().ndb_b200_knn_regress(self.h, ptr(t), ptr(Q), Q.shape[0], k, ptr(out)))
        return out

    def knn_dev(self, q_ptr, nq, k, dist_ptr, ids_ptr, metric=L2, arith=ARITH_OP_F64, stream=None):
        check(L.load().ndb_b200_knn_exact_dev(self.h, metric, arith, ptr(q_ptr), nq, k, ptr(dist_ptr), ptr(ids_ptr),
                                              ptr(stream)))

    def knn_sharded_dev(self, q_ptr, nq, k, dist_ptr, ids_ptr, metric=L2, arith=ARITH_OP_F64, stream=None):
        """Rows split between the ranks (this handle holds this rank's, with global ids), queries replicated:
        local top-k, one all-gather, (dist, id) merge -- every rank gets the full result."""
        check(L.load().ndb_b200_knn_exact_sharded_dev(self.h, metric, arith, ptr(q_ptr), nq, k, ptr(dist_ptr),
                                                      ptr(ids_ptr), ptr(stream)))


class IvfIndex(_Handle):
    """CREATE INDEX ... USING ivf (v vector_l2_ops) WITH (lists = L)  (ivf_am.c)."""
    _free = "ndb_b200_ivf_free"

    def __init__(self, dim, lists=100, metric=L2):
        super().__init__()
        self.dim, self.lists, self.metric = dim, lists, metric
        check(L.load().ndb_b200_ivf_create(dim, lists, metric, C.byref(self.h)))

    def set_shard(self, rank, world):
        check(L.load().ndb_b200_ivf_set_shard(self.h, rank, world))

    def ivfbuild(self, rows):
        """ivfbuild (:501-745): k-means on the first min(10000, lists*100) rows; no list assignment (Q6)."""
        rows = _rows(rows, self.dim, "ivfbuild")
        check(L.load().ndb_b200_ivf_train(self.h, ptr(rows), rows.shape[0]))

    def set_centroids(self, Cn):
        Cn = f32(Cn)
        assert Cn.shape == (self.lists, self.dim)
        check(L.load().ndb_b200_ivf_set_centroids(self.h, ptr(Cn)))

    def centroids(self):
        Cn = np.empty((self.lists, self.dim), np.float32)
        check(L.load().ndb_b200_ivf_get_centroids(self.h, ptr(Cn)))
        return Cn

    def ivfinsert(self, rows, ids=None):
        """ivfinsert (:797-1167) for each row in order; returns the list id of every row."""
        rows = _rows(rows, self.dim, "ivfinsert")
        idp = None if ids is None else np.ascontiguousarray(ids, np.int64)
        out = np.empty(rows.shape[0], np.int32)
        check(L.load().ndb_b200_ivf_insert(self.h, ptr(rows), ptr(idp), rows.shape[0], ptr(out)))
        return out

    def assign(self, rows):
        rows = _rows(rows, self.dim, "ivf_assign")
        out = np.empty(rows.shape[0], np.int32)
        check(L.load().ndb_b200_ivf_assign(self.h, ptr(rows), rows.shape[0], ptr(out)))
        return out

    def prepare(self, arith=ARITH_IVF_F32):
        """List layout (+ blocked bf16 copy for ARITH_TENSOR) now instead of inside the first search."""
        check(L.load().ndb_b200_ivf_prepare(self.h, arith))

    def cert_stats(self):
        """Certified selection of the last ARITH_TENSOR search: dict of exact-fallback queries and exact
        re-evaluations, for the list scan and for the coarse quantiser."""
        out = np.zeros(6, np.int64)
        check(L.load().ndb_b200_ivf_cert_stats(self.h, ptr(out)))
        return {"list_fallback_queries": int(out[0]), "list_exact_evals": int(out[1]),
                "coarse_fallback_queries": int(out[2]), "coarse_exact_evals": int(out[3]),
                "list_full_scan_queries": int(out[4]), "list_rescanned_rows": int(out[5])}

    def load_relation(self, blocks):
        blocks = np.ascontiguousarray(blocks, np.uint8)
        check(L.load().ndb_b200_ivf_load_relation(self.h, ptr(blocks), blocks.size // 8192))

    def __len__(self):
        return int(L.load().ndb_b200_ivf_size(self.h))

    def list_sizes(self):
        out = np.empty(self.lists, np.int64)
        check(L.load().ndb_b200_ivf_list_sizes(self.h, ptr(out)))
        return out

    def select_clusters(self, Q, nprobe):
        Q = _rows(Q, self.dim, "ivf_select_clusters")
        out = np.empty((Q.shape[0], nprobe), np.int32)
        check(L.load().ndb_b200_ivf_select_clusters(self.h, ptr(Q), Q.shape[0], nprobe, ptr(out)))
        return out

    def search(self, Q, nprobe=10, k=10, mode=IVF_FULL, arith=ARITH_IVF_F32):
        """ivfrescan + ivfgettuple for a batch of queries (:1439-1545, 1911-2027)."""
        Q = _rows(Q, self.dim, "ivf_search")
        d = np.empty((Q.shape[0], k), np.float32)
        i = np.empty((Q.shape[0], k), np.int64)
        check(L.load().ndb_b200_ivf_search(self.h, ptr(Q), Q.shape[0], nprobe, k, mode, arith, ptr(d), ptr(i)))
        return d, i

    def search_begin(self, Q, dist, ids, nprobe=10, k=10, mode=IVF_FULL, arith=ARITH_IVF_F32):
        """Queue one batch (Q [nq, dim] float32, dist [nq, k] float32, ids [nq, k] int64: numpy arrays the
        caller keeps alive, ideally pinned); returns a ticket for search_end.  Two batches may be in flight."""
        _pinned_like(Q, np.float32, (Q.shape[0] if getattr(Q, "ndim", 0) == 2 else -1, self.dim), "ivf_search_begin Q")
        _pinned_like(dist, np.float32, (Q.shape[0], k), "ivf_search_begin dist")
        _pinned_like(ids, np.int64, (Q.shape[0], k), "ivf_search_begin ids")
        t = C.c_int(-1)
        check(L.load().ndb_b200_ivf_search_begin(self.h, ptr(Q), Q.shape[0], nprobe, k, mode, arith, ptr(dist), ptr(ids),
                                                 C.byref(t)))
        return t.value

    def search_end(self, ticket):
        check(L.load().ndb_b200_ivf_search_end(self.h, ticket))

    def search_dev(self, q_ptr, nq, dist_ptr, ids_ptr, nprobe=10, k=10, mode=IVF_FULL, arith=ARITH_IVF_F32,
                   stream=None):
        check(L.load().ndb_b200_ivf_search_dev(self.h, ptr(q_ptr), nq, nprobe, k, mode, arith, ptr(dist_ptr),
                                               ptr(ids_ptr), ptr(stream)))


    def knn_search_gpu(self, query, k, nprobe=10):
        """ivf_knn_search_gpu(index, query, k, nprobe) (gpu_sql.c:931): (ids, distances) of the rows found."""
        q = f32(query).reshape(-1)
        ids = np.empty(max(k, 1), np.int64)
        d = np.empty(max(k, 1), np.float32)
        nres = C.c_int()
        check(L.load().ndb_b200_ivf_knn_search_gpu(self.h, ptr(q), q.shape[0], k, nprobe, ptr(ids), ptr(d), C.byref(nres)))
        return ids[:nres.value], d[:nres.value]

    def search_sharded_dev(self, q_ptr, nq, dist_ptr, ids_ptr, nprobe=10, k=10, mode=IVF_FULL, arith=ARITH_IVF_F32,
                           stream=None):
        """Local search + all-gather of packed (dist, id) records + device merge (ndb_b200_ivf_search_sharded_dev)."""
        check(L.load().ndb_b200_ivf_search_sharded_dev(self.h, ptr(q_ptr), nq, nprobe, k, mode, arith, ptr(dist_ptr),
                                                       ptr(ids_ptr), ptr(stream)))

    def search_sharded(self, Q, nprobe=10, k=10, mode=IVF_FULL, arith=ARITH_IVF_F32, dist=None, ids=None):
        Q = _rows(Q, self.dim, "ivf_search_sharded")
        d = np.empty((Q.shape[0], k), np.float32) if dist is None else dist
        i = np.empty((Q.shape[0], k), np.int64) if ids is None else ids
        check(L.load().ndb_b200_ivf_search_sharded(self.h, ptr(Q), Q.shape[0], nprobe, k, mode, arith, ptr(d), ptr(i)))
        return d, i


class HnswIndex(_Handle):
    """CREATE INDEX ... USING hnsw (v ...) WITH (m, ef_construction, ef_search)  (hnsw_am.c)."""
    _free = "ndb_b200_hnsw_free"

    def __init__(self, dim, m=16, ef_construction=200, ef_search=64, metric=L2):
        super().__init__()
        self.dim, self.m, self.efc, self.efs, self.metric = dim, m, ef_construction, ef_search, metric
        check(L.load().ndb_b200_hnsw_create(dim, m, ef_construction, ef_search, metric, C.byref(self.h)))

    def hnswbuild(self, rows, ids=None, levels=None, seed=0, batch=0, select=HNSW_SELECT_CLOSEST):
        """hnswbuild (:343-415).  select=HNSW_SELECT_HEURISTIC swaps the reference's closest-m / capped
        back-link rule for the diversity heuristic with re-selection of full neighbours (an extension)."""
        check(L.load().ndb_b200_hnsw_set_select(self.h, select))
        rows = _rows(rows, self.dim, "hnswbuild")
        idp = None if ids is None else np.ascontiguousarray(ids, np.int64)
        lv = None if levels is None else np.ascontiguousarray(levels, np.int32)
        check(L.load().ndb_b200_hnsw_build(self.h, ptr(rows), ptr(idp), rows.shape[0], ptr(lv), seed, batch))

    def load_graph(self, rows, g, ids=None):
        rows = f32(rows)
        idp = None if ids is None else np.ascontiguousarray(ids, np.int64)
        self._keep = [np.ascontiguousarray(g["levels"], np.int32), np.ascontiguousarray(g["nbr0"], np.uint32),
                      np.ascontiguousarray(g["cnt"], np.int16), np.ascontiguousarray(g["upper_off"], np.int64),
                      np.ascontiguousarray(g["upper"], np.uint32)]
        lv, n0, cn, uo, up = self._keep
        check(L.load().ndb_b200_hnsw_load_graph(self.h, ptr(rows), ptr(idp), rows.shape[0], ptr(lv), ptr(n0), ptr(cn),
                                                ptr(uo), ptr(up), int(g["entry_point"]), int(g["entry_level"])))

    def export_graph(self):
        n = len(self)
        levels = np.empty(n, np.int32)
        nbr0 = np.empty((n, 2 * self.m), np.uint32)
        cnt = np.empty((n, 16), np.int16)
        uoff = np.empty(n + 1, np.int64)
        cap = max(1, n * 2 * self.m)        # generous: sum(level) * 2m <= n * 2m for ml = 0.36
        upper = np.empty(cap, np.uint32)
        ep, el = C.c_uint32(), C.c_int()
        check(L.load().ndb_b200_hnsw_export_graph(self.h, ptr(levels), ptr(nbr0), ptr(cnt), ptr(uoff), ptr(upper), cap,
                                                  C.byref(ep), C.byref(el)))
        return dict(levels=levels, nbr0=nbr0, cnt=cnt, upper_off=uoff, upper=upper[:max(1, int(uoff[n]))],
                    entry_point=ep.value, entry_level=el.value)

    def load_relation(self, blocks):
        blocks = np.ascontiguousarray(blocks, np.uint8)
        check(L.load().ndb_b200_hnsw_load_relation(self.h, ptr(blocks), blocks.size // 8192))

    def __len__(self):
        return int(L.load().ndb_b200_hnsw_size(self.h))

    def search(self, Q, ef=None, k=10, strategy=None, mode=HNSW_BESTFIRST):
        """hnswrescan + hnswgettuple for a batch (:904-1056); ef defaults to the index's ef_search."""
        Q = _rows(Q, self.dim, "hnsw_search")
        d = np.empty((Q.shape[0], k), np.float32)
        i = np.empty((Q.shape[0], k), np.int64)
        check(L.load().ndb_b200_hnsw_search(self.h, ptr(Q), Q.shape[0], strategy or self.metric, ef or self.efs, k,
                                            mode, ptr(d), ptr(i)))
        return d, i

    def search_dev(self, q_ptr, nq, dist_ptr, ids_ptr, ef=None, k=10, strategy=None, mode=HNSW_BESTFIRST, stream=None):
        check(L.load().ndb_b200_hnsw_search_dev(self.h, ptr(q_ptr), nq, strategy or self.metric, ef or self.efs, k, mode,
                                                ptr(dist_ptr), ptr(ids_ptr), ptr(stream)))

    def last_evals(self):
        return int(L.load().ndb_b200_hnsw_last_evals(self.h))

    def knn_search_gpu(self, query, k, ef_search=100):
        """hnsw_knn_search_gpu(index, query, k, ef_search) (gpu_sql.c:498): (ids, distances) of the rows found."""
        q = f32(query).reshape(-1)
        ids = np.empty(max(k, 1), np.int64)
        d = np.empty(max(k, 1), np.float32)
        nres = C.c_int()
        check(L.load().ndb_b200_hnsw_knn_search_gpu(self.h, ptr(q), q.shape[0], k, ef_search, ptr(ids), ptr(d), C.byref(nres)))
        return ids[:nres.value], d[:nres.value]

    def broadcast(self, n=None, root=0):
        """Replicas: the rank that built the graph sends it to the others (collective)."""
        check(L.load().ndb_b200_hnsw_broadcast(self.h, root))
