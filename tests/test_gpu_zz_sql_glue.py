"""GPU: the SQL-function glue (integration/ml_sql_b200.c: cluster_kmeans, cluster_minibatch_kmeans, train_pq_codebook,
pq_encode_vector as the reference's functions would call them, rows as float **, rand() drawn by the glue) against the
direct ABI calls with the same rand() values.  (Named to run after the other GPU files.)"""
import ctypes as C
import os

import numpy as np
import pytest

import workloads as W

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BITS = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def glue(ndb):
    path = os.path.join(ROOT, "oracle", "_ref", "libndb_b200_glue.so")
    assert os.path.exists(path), "oracle/_ref/libndb_b200_glue.so must travel with the snapshot (make glue)"
    g = C.CDLL(path)
    assert g.ndb_b200_glue_init() == 0
    return g


def test_sql_function_glue_equals_the_direct_calls(ndb, orc, glue):
    """integration/ml_sql_b200.c on the GPU: the glue draws from the process's rand() as the reference does; with the
    generator seeded alike, its results are those of the ABI calls given the same values (which the oracle pins)."""
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    libc = C.CDLL(None)
    X = W.mixture(3000, 16, 6, 41)
    # cluster_kmeans: k draws
    draws = orc.libc_rand_draws(5, 6)
    want_l, _, _, _ = ndb.cluster_kmeans(X, 6, 4, draws)
    libc.srand(5)
    labels = np.zeros(3000, np.int32)
    assert glue.ndb_b200_glue_cluster_kmeans(p(X), 3000, 16, 6, 4, p(labels)) == 0
    assert np.array_equal(labels, want_l) and np.array_equal(labels, orc.cluster_kmeans(X, 6, 4, draws)[0])
    # cluster_minibatch_kmeans: the generator itself
    draws = orc.libc_rand_draws(6, 6 + 50 * 8)
    want_l, _, used = ndb.cluster_minibatch_kmeans(X, 6, 50, 8, draws)
    libc.srand(6)
    assert glue.ndb_b200_glue_cluster_minibatch_kmeans(p(X), 3000, 16, 6, 50, 8, p(labels)) == 0
    assert np.array_equal(labels, want_l) and used == 6 + 50 * 8
    # train_pq_codebook (100 iterations) + pq_encode_vector on the bytea payload
    draws = orc.libc_rand_draws(7, 4 * 16)
    want_cb = ndb.pq_train(X, 4, 16, draws, 100)
    libc.srand(7)
    cb = np.zeros((4, 16, 4), np.float32)
    assert glue.ndb_b200_glue_train_pq_codebook(p(X), 3000, 16, 4, 16, p(cb)) == 0
    assert np.array_equal(BITS(cb), BITS(want_cb))
    payload = np.array([4, 16, 4], np.int32).tobytes() + cb.tobytes()
    codes = np.zeros(4, np.int16)
    assert glue.ndb_b200_glue_pq_encode_vector(p(X[11:12]), 16, payload, p(codes)) == 0
    assert np.array_equal(codes, orc.pq_encode(X[11:12], cb)[0])
