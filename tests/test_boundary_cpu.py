"""CPU: the reference-side glue (integration/*.c) is compiled against the reference's own
neurondb_gpu_backend.h -- a signature mismatch fails the build -- and the resulting library exposes the
backend instance the registry would select.  No compute calls without a GPU."""
import ctypes as C
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLUE = os.path.join(ROOT, "oracle", "_ref", "libndb_b200_glue.so")
REF_HEADER = "/root/reference/NeuronDB/include/neurondb_gpu_backend.h"


def glue():
    if not os.path.exists(GLUE):
        pytest.skip("oracle/_ref/libndb_b200_glue.so not built (no reference tree and no prebuilt copy)")
    return C.CDLL(GLUE)


def test_glue_builds_against_the_reference_header():
    if not os.path.exists(REF_HEADER):
        pytest.skip("reference tree absent")
    out = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "glue"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]
    assert os.path.exists(GLUE)


def test_backend_instance_identity():
    g = glue()
    g.ndb_b200_glue_name.restype = C.c_char_p
    assert g.ndb_b200_glue_name() == b"b200"
    assert g.ndb_b200_glue_priority() > 90          # above the CUDA backend's .priority = 90 (gpu_backend_cuda.c:734-740)
    assert g.ndb_b200_glue_unsupported_members_are_null() == 1
    for sym in ("neurondb_gpu_b200_backend", "ndb_b200_am_stage_ivf", "ndb_b200_am_stage_hnsw", "ndb_b200_am_ivf_beginscan",
                "ndb_b200_am_hnsw_beginscan", "ndb_b200_am_rescan", "ndb_b200_am_gettuple", "ndb_b200_am_endscan"):
        assert hasattr(g, sym), sym


def test_without_a_gpu_the_backend_fails_loudly():
    """No CPU fallback: on a host without a device init() and the launchers return a negative code."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    g = glue()
    assert g.ndb_b200_glue_init() < 0
    a = (C.c_float * 4)(1, 2, 3, 4)
    out = (C.c_float * 1)()
    assert g.ndb_b200_glue_l2(a, a, out, 1, 4) < 0


def test_sql_function_glue_helpers():
    """integration/ml_sql_b200.c: the float ** rows of neurondb_fetch_vectors_from_table flatten to the array the ABI takes,
    the draws are the backend's rand() values in order, and without a device the functions fail loudly."""
    import numpy as np
    g = glue()
    for sym in ("ndb_b200_sql_cluster_kmeans", "ndb_b200_sql_cluster_minibatch_kmeans", "ndb_b200_sql_train_pq_codebook",
                "ndb_b200_sql_pq_encode_vector", "ndb_b200_sql_flatten_rows", "ndb_b200_sql_draw"):
        assert hasattr(g, sym), sym
    X = np.random.default_rng(3).standard_normal((37, 5)).astype(np.float32)
    assert g.ndb_b200_glue_flatten_check(X.ctypes.data_as(C.c_void_p), 37, 5) == 1
    libc = C.CDLL(None)
    libc.srand(77)
    libc.rand.restype = C.c_int
    want = [libc.rand() for _ in range(9)]
    libc.srand(77)
    got = (C.c_int * 9)()
    g.ndb_b200_glue_draw(9, got)
    assert list(got) == want
    import torch
    if torch.cuda.is_available():
        return
    labels = (C.c_int * 37)()
    assert g.ndb_b200_glue_cluster_kmeans(X.ctypes.data_as(C.c_void_p), 37, 5, 3, 4, labels) == -3          # ENOTINIT
    assert g.ndb_b200_glue_cluster_minibatch_kmeans(X.ctypes.data_as(C.c_void_p), 37, 5, 3, 8, 4, labels) == -3
    cb = np.zeros((1, 4, 5), np.float32)
    assert g.ndb_b200_glue_train_pq_codebook(X.ctypes.data_as(C.c_void_p), 37, 5, 1, 4, cb.ctypes.data_as(C.c_void_p)) == -3
    assert g.ndb_b200_glue_train_pq_codebook(X.ctypes.data_as(C.c_void_p), 37, 5, 2, 4, cb.ctypes.data_as(C.c_void_p)) == -1   # 5 % 2
    payload = np.array([1, 4, 4], np.int32).tobytes() + cb.tobytes()                                           # dsub 4 != dim 5
    codes = (C.c_int16 * 1)()
    assert g.ndb_b200_glue_pq_encode_vector(X.ctypes.data_as(C.c_void_p), 5, payload, codes) == -5           # EDIM
