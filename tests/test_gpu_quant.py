"""GPU: the per-vector quantisers of NeuronDB/src/types/quantization.c for row sets (quantize_vector_i8 / _f16 / _binary /
_uint8 / _ternary / _int4) and ORDER BY binary_hamming_distance LIMIT k, through the C ABI.

Every output byte must equal (a) the committed outputs of the reference's OWN functions (tests/golden/ml_paths.npz,
written from oracle/_ref/libndb_ref_leafs.so by tests/golden/make_golden.py) and (b) the oracle restatement on larger
inputs."""
import os

import numpy as np
import pytest

import oracle_lib as O
import workloads as W
from test_oracle import _quant_inputs

pytestmark = pytest.mark.gpu
KINDS = (O.Q_INT8, O.Q_FP16, O.Q_BINARY, O.Q_UINT8, O.Q_TERNARY, O.Q_INT4)
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ml_paths.npz")


def test_quantisers_equal_the_reference_outputs(ndb):
    g = np.load(GOLDEN)
    for X in _quant_inputs():
        for kind in KINDS:
            assert np.array_equal(ndb.quantize_rows(kind, X), g["quant_k%d_d%d" % (kind, X.shape[1])]), (kind, X.shape)


@pytest.mark.parametrize("n,dim", [(20000, 128), (3000, 97), (1000, 768), (5, 1)])
def test_quantisers_equal_the_oracle(ndb, n, dim):
    rng = np.random.default_rng(n + dim)
    X = W.mixture(n, dim, 9, n + dim) * rng.choice([1e-5, 1.0, 7e4], (n, 1)).astype(np.float32)
    X[rng.integers(0, n, max(1, n // 20))] = 0.0
    X[rng.integers(0, n)] = 3.25
    X = np.where(rng.random(X.shape) < 0.05, np.round(X), X).astype(np.float32)          # exact .0 / ties for rintf
    for kind in KINDS:
        got = ndb.quantize_rows(kind, X)
        assert got.shape == (n, O.lib().orc_quantized_row_bytes(kind, dim))
        assert np.array_equal(got, O.quantize_rows(kind, X)), kind


@pytest.mark.parametrize("n,nbits,nq,k", [(50000, 128, 37, 10), (7000, 100, 9, 100), (3000, 1024, 16, 32), (20, 9, 3, 30)])
def test_hamming_scan_equals_the_oracle(ndb, n, nbits, nq, k):
    X = W.gaussian(n, nbits, n)
    Q = W.gaussian(nq, nbits, n + 1)
    rows, qs = ndb.quantize_rows(ndb.QUANT_BINARY, X), ndb.quantize_rows(ndb.QUANT_BINARY, Q)
    rows[n // 2] = qs[0]                                                              # an exact match: distance 0
    wd, wi = O.hamming_knn(rows, nbits, qs, k)
    d, i = ndb.hamming_knn(rows, nbits, qs, k)
    assert np.array_equal(d, wd) and np.array_equal(i, wi)                            # ties by row, -1 / -1 past the end
    assert d[0, 0] == 0 and i[0, 0] <= n // 2
    os.environ["NDB_HAMMING_PER_QUERY"] = "1"                                         # the warp-per-query kernel gives the same
    try:
        d2, i2 = ndb.hamming_knn(rows, nbits, qs, k)
    finally:
        del os.environ["NDB_HAMMING_PER_QUERY"]
    assert np.array_equal(d2, wd) and np.array_equal(i2, wi)


def test_quantiser_errors(ndb):
    X = W.gaussian(4, 8, 1)
    with pytest.raises(ndb.NdbError) as e:
        ndb.quantize_rows(9, X)
    assert e.value.code == -1
    bad = X.copy()
    bad[2, 3] = np.nan
    with pytest.raises(ndb.NdbError) as e:
        ndb.quantize_rows(ndb.QUANT_UINT8, bad)
    assert e.value.code == -4
    with pytest.raises(ndb.NdbError) as e:
        ndb.hamming_knn(np.zeros((4, 2), np.uint8), 16, np.zeros((1, 3), np.uint8), 2)
    assert e.value.code == -5 and "binary vector dimensions must match" in str(e.value)
