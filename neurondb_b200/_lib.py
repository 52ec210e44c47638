"""ctypes loader for libndb_b200.so -- the C-ABI declared in include/ndb_b200.h.

The library is the product; this module only binds it.  There is no fallback: if the shared
object is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NDB_B200_LIB_PATH") or os.path.join(_HERE, "lib", "libndb_b200.so")   # (override: instrumented development builds)

OK = 0
L2, COSINE, IP = 1, 2, 3
ARITH_OP_F64, ARITH_IVF_F32, ARITH_HNSW, ARITH_FAST, ARITH_TENSOR = 0, 3, 4, 5, 6
ARITH_AVX2, ARITH_AVX512 = 1, 2
IVF_FULL, IVF_LITERAL = 0, 1
QUANT_INT8, QUANT_FP16, QUANT_BINARY, QUANT_UINT8, QUANT_TERNARY, QUANT_INT4 = 1, 2, 3, 4, 5, 6
HNSW_LITERAL, HNSW_BESTFIRST = 0, 1
HNSW_SELECT_CLOSEST, HNSW_SELECT_HEURISTIC = 0, 1

ERRORS = {-1: "EINVAL", -2: "ECUDA", -3: "ENOTINIT", -4: "EVECTOR", -5: "EDIM", -6: "ENOMEM", -7: "ESTATE",
          -8: "ERANGE"}


class NdbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ndb_b200 error %d (%s): %s" % (code, ERRORS.get(code, "?"), msg))
        self.code = code


_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_u32 = C.c_uint32
_f = C.c_float
_sz = C.c_size_t

# name -> (restype, argtypes); the list mirrors include/ndb_b200.h one to one
SIGNATURES = {
    "ndb_b200_init": (_i, [_i]),
    "ndb_b200_shutdown": (None, []),
    "ndb_b200_is_available": (_i, []),
    "ndb_b200_device_count": (_i, []),
    "ndb_b200_abi_version": (_i, []),
    "ndb_b200_last_error": (C.c_char_p, []),
    "ndb_b200_device_info": (_i, [_i, C.c_char_p, _sz, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_i), C.POINTER(_i),
                                  C.POINTER(_i)]),
    "ndb_b200_mem_alloc": (_i, [C.POINTER(_p), _sz]),
    "ndb_b200_mem_free": (_i, [_p]),
    "ndb_b200_memcpy_h2d": (_i, [_p, _p, _sz]),
    "ndb_b200_memcpy_d2h": (_i, [_p, _p, _sz]),
    "ndb_b200_host_alloc_pinned": (_i, [C.POINTER(_p), _sz]),
    "ndb_b200_host_free_pinned": (_i, [_p]),
    "ndb_b200_stream_create": (_i, [C.POINTER(_p)]),
    "ndb_b200_stream_destroy": (_i, [_p]),
    "ndb_b200_stream_synchronize": (_i, [_p]),
    "ndb_b200_launch_l2_distance": (_i, [_p, _p, _p, _i, _i, _p]),
    "ndb_b200_launch_cosine": (_i, [_p, _p, _p, _i, _i, _p]),
    "ndb_b200_launch_kmeans_assign": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "ndb_b200_launch_kmeans_update": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "ndb_b200_distance_pairs": (_i, [_i, _i, _p, _p, _p, _i64, _i]),
    "ndb_b200_distance_rows": (_i, [_i, _i, _p, _i64, _i, _p, _p]),
    "ndb_b200_dataset_create": (_i, [_i, C.POINTER(_p)]),
    "ndb_b200_dataset_append": (_i, [_p, _p, _p, _i64]),
    "ndb_b200_dataset_append_dev": (_i, [_p, _p, _p, _i64, _p]),
    "ndb_b200_dataset_size": (_i64, [_p]),
    "ndb_b200_dataset_free": (None, [_p]),
    "ndb_b200_knn_exact": (_i, [_p, _i, _i, _p, _i, _i, _p, _p]),
    "ndb_b200_knn_exact_dev": (_i, [_p, _i, _i, _p, _i, _i, _p, _p, _p]),
    "ndb_b200_knn_classify": (_i, [_p, _p, _p, _i, _i, _p]),
    "ndb_b200_knn_regress": (_i, [_p, _p, _p, _i, _i, _p]),
    "ndb_b200_cluster_kmeans": (_i, [_p, _i, _i, _i, _i, _p, _i, _p, _p, _p, _p]),
    "ndb_b200_cluster_minibatch_kmeans": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _i, _p, _p]),
    "ndb_b200_quantized_row_bytes": (_i64, [_i, _i]),
    "ndb_b200_quantize_rows": (_i, [_i, _p, _i64, _i, _p]),
    "ndb_b200_hamming_knn": (_i, [_p, _i64, _i, _p, _i, _i, _p, _p]),
    "ndb_b200_pq_train": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "ndb_b200_pq_encode": (_i, [_p, _i64, _i, _p, _i, _i, _p]),
    "ndb_b200_launch_pq_encode": (_i, [_p, _p, _p, _i, _i, _i, _i, _p]),
    "ndb_b200_pq_create": (_i, [_i, _i, _i, _p, C.POINTER(_p)]),
    "ndb_b200_pq_free": (None, [_p]),
    "ndb_b200_pq_size": (_i64, [_p]),
    "ndb_b200_pq_add": (_i, [_p, _p, _i64, _p]),
    "ndb_b200_pq_add_codes": (_i, [_p, _p, _i64]),
    "ndb_b200_pq_search": (_i, [_p, _p, _i, _i, _p, _p]),
    "ndb_b200_pq_search_dev": (_i, [_p, _p, _i, _i, _p, _p, _p]),
    "ndb_b200_pq_distances": (_i, [_p, _p, _i, _p, _p]),
    "ndb_b200_kmeans_train": (_i, [_p, _i, _i, _i, _i, _f, _p, _p, _p, C.POINTER(_i), C.POINTER(_f)]),
    "ndb_b200_ivf_create": (_i, [_i, _i, _i, C.POINTER(_p)]),
    "ndb_b200_ivf_free": (None, [_p]),
    "ndb_b200_ivf_train": (_i, [_p, _p, _i64]),
    "ndb_b200_ivf_set_centroids": (_i, [_p, _p]),
    "ndb_b200_ivf_get_centroids": (_i, [_p, _p]),
    "ndb_b200_ivf_insert": (_i, [_p, _p, _p, _i64, _p]),
    "ndb_b200_ivf_assign": (_i, [_p, _p, _i64, _p]),
    "ndb_b200_ivf_load_relation": (_i, [_p, _p, _u32]),
    "ndb_b200_ivf_size": (_i64, [_p]),
    "ndb_b200_ivf_list_sizes": (_i, [_p, _p]),
    "ndb_b200_ivf_search": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "ndb_b200_ivf_search_dev": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "ndb_b200_ivf_select_clusters": (_i, [_p, _p, _i, _i, _p]),
    "ndb_b200_ivf_set_shard": (_i, [_p, _i, _i]),
    "ndb_b200_hnsw_create": (_i, [_i, _i, _i, _i, _i, C.POINTER(_p)]),
    "ndb_b200_hnsw_free": (None, [_p]),
    "ndb_b200_hnsw_build": (_i, [_p, _p, _p, _i64, _p, C.c_uint, _i]),
    "ndb_b200_hnsw_load_graph": (_i, [_p, _p, _p, _i64, _p, _p, _p, _p, _p, _u32, _i]),
    "ndb_b200_hnsw_export_graph": (_i, [_p, _p, _p, _p, _p, _p, _i64, C.POINTER(_u32), C.POINTER(_i)]),
    "ndb_b200_hnsw_size": (_i64, [_p]),
    "ndb_b200_hnsw_load_relation": (_i, [_p, _p, _u32]),
    "ndb_b200_hnsw_search": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "ndb_b200_hnsw_search_dev": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "ndb_b200_hnsw_last_evals": (_i64, [_p]),
    "ndb_b200_merge_topk_dev": (_i, [_p, _p, _i, _i, _i, _p, _p, _p]),
    "ndb_b200_merge_topk": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "ndb_b200_kmeans_shard_step_dev": (_i, [_p, _i64, _i, _i, _p, _p, _p, _p, _p]),
    "ndb_b200_kmeans_shard_cost_dev": (_i, [_p, _i64, _i, _p, _p, _p, _p]),
    "ndb_b200_ivf_search_begin": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "ndb_b200_ivf_search_end": (_i, [_p, _i]),
    "ndb_b200_hnsw_set_select": (_i, [_p, _i]),
    "ndb_b200_keys_from_vector": (_i, [_p, _i64, _i, _p]),
    "ndb_b200_keys_from_halfvec": (_i, [_p, _i64, _i, _p]),
    "ndb_b200_keys_from_halfvec_dev": (_i, [_p, _i64, _i, _p, _p]),
    "ndb_b200_keys_from_bits": (_i, [_p, _i64, _i, _p]),
    "ndb_b200_keys_from_bits_dev": (_i, [_p, _i64, _i, _p, _p]),
    "ndb_b200_keys_from_sparse": (_i, [_p, _p, _p, _i64, _i, _p]),
    "ndb_b200_vector_distance_batch": (_i, [_i, _p, _p, _i64, _i, _p, _p, _p]),
    "ndb_b200_ivf_knn_search_gpu": (_i, [_p, _p, _i, _i, _i, _p, _p, C.POINTER(_i)]),
    "ndb_b200_hnsw_knn_search_gpu": (_i, [_p, _p, _i, _i, _i, _p, _p, C.POINTER(_i)]),
    "ndb_b200_ivf_dim": (_i, [_p]),
    "ndb_b200_ivf_prepare": (_i, [_p, _i]),
    "ndb_b200_ivf_cert_stats": (_i, [_p, _p]),
    "ndb_b200_hnsw_broadcast": (_i, [_p, _i]),
    "ndb_b200_comm_unique_id": (_i, [_p, _sz]),
    "ndb_b200_comm_init": (_i, [_i, _i, _p, _sz]),
    "ndb_b200_comm_shutdown": (_i, []),
    "ndb_b200_comm_rank": (_i, []),
    "ndb_b200_comm_nranks": (_i, []),
    "ndb_b200_comm_nccl_version": (_i, []),
    "ndb_b200_comm_exchange_is_p2p": (_i, []),
    "ndb_b200_comm_allgather_dev": (_i, [_p, _p, _sz, _p]),
    "ndb_b200_comm_allreduce_sum_dev": (_i, [_p, _sz, _i, _p]),
    "ndb_b200_comm_broadcast_dev": (_i, [_p, _sz, _i, _p]),
    "ndb_b200_ivf_search_sharded_dev": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "ndb_b200_ivf_search_sharded": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "ndb_b200_knn_exact_sharded_dev": (_i, [_p, _i, _i, _p, _i, _i, _p, _p, _p]),
    "ndb_b200_kmeans_train_sharded_dev": (_i, [_p, _i64, _i, _i, _i, _f, _p, _p, _p, C.POINTER(_i), C.POINTER(_f), _p]),
    "ndb_b200_launch_count": (_i64, []),
    "ndb_b200_set_timing": (_i, [_i]),
    "ndb_b200_last_kernel_stats": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i64)]),
}

_lib = None


def load():
    """Load the shared library (no CUDA call is made here)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libndb_b200.so is not built (%s). Run `make` or `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        raise NdbError(rc, (load().ndb_b200_last_error() or b"").decode(errors="replace"))
    return rc


def ptr(a):
    """Host numpy array -> void*, or an int device pointer passed through."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)
