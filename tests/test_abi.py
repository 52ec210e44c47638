"""CPU: the C-ABI library loads and exports every symbol include/ndb_b200.h declares."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ndb_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ndb_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    fns = declared_functions()
    for must in ("ndb_b200_init", "ndb_b200_launch_l2_distance", "ndb_b200_launch_cosine",
                 "ndb_b200_launch_kmeans_assign", "ndb_b200_launch_kmeans_update", "ndb_b200_distance_pairs",
                 "ndb_b200_knn_exact", "ndb_b200_kmeans_train", "ndb_b200_ivf_train", "ndb_b200_ivf_insert",
                 "ndb_b200_ivf_search", "ndb_b200_hnsw_build", "ndb_b200_hnsw_search", "ndb_b200_merge_topk_dev"):
        assert must in fns


def test_library_exports_every_declared_symbol():
    import neurondb_b200._lib as L
    assert os.path.exists(L.LIB_PATH), "libndb_b200.so is not built: run make"
    out = subprocess.check_output(["nm", "-D", "--defined-only", L.LIB_PATH], text=True)
    exported = set(re.findall(r" T (ndb_b200_[a-z0-9_]+)", out))
    missing = [f for f in declared_functions() if f not in exported]
    assert not missing, missing
    # and nothing is exported that the header does not declare
    assert not (exported - set(declared_functions())), exported - set(declared_functions())


def test_ctypes_binding_covers_the_header():
    import neurondb_b200._lib as L
    assert sorted(L.SIGNATURES) == declared_functions()
    lib = L.load()
    assert lib.ndb_b200_abi_version() == 2


def test_no_cpu_fallback_without_a_device():
    """On a box without a GPU the library must fail loudly, not compute on the CPU."""
    import neurondb_b200 as ndb
    lib = ndb._lib.load()
    if lib.ndb_b200_device_count() > 0:
        pytest.skip("a CUDA device is visible here")
    assert not ndb.is_available()
    with pytest.raises(ndb.NdbError) as e:
        ndb.init(0)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)
    with pytest.raises(ndb.NdbError) as e:
        ndb.distance_pairs([[1.0, 2.0]], [[2.0, 3.0]])
    assert e.value.code == -3


def test_next_rows_fail_loudly_without_a_device_and_check_their_arguments_first():
    """The SURVEY 8f entry points (cluster_kmeans, minibatch, PQ, quantisers, Hamming scan): ENOTINIT without a device, and
    the host-side mirror rejects malformed arrays before anything reaches the library."""
    import numpy as np
    import neurondb_b200 as ndb
    lib = ndb._lib.load()
    X = np.zeros((8, 4), np.float32)
    d = np.arange(16, dtype=np.int32)
    # argument shapes: caught by the mirror, device or not
    for call in (lambda: ndb.cluster_kmeans(X, 3, 2, d[:2]),                       # k draws expected
                 lambda: ndb.cluster_kmeans(X[0], 2, 2, d[:2]),                    # rows are 2-d
                 lambda: ndb.pq_train(X, 2, 4, d[:3]),                             # m * ksub draws expected
                 lambda: ndb.pq_encode(X, np.zeros((2, 4), np.float32)),           # codebooks are [m][ksub][dsub]
                 lambda: ndb.quantize_rows(ndb.QUANT_INT8, X[0])):
        with pytest.raises(ndb.NdbError) as e:
            call()
        assert e.value.code == -1
    with pytest.raises(ndb.NdbError) as e:
        ndb.pq_encode(X, np.zeros((2, 4, 3), np.float32))                          # 2 * 3 != 4 columns
    assert e.value.code == -5
    with pytest.raises(ndb.NdbError) as e:
        ndb.hamming_knn(np.zeros((4, 2), np.uint8), 16, np.zeros((1, 3), np.uint8), 2)
    assert e.value.code == -5 and "binary vector dimensions must match" in str(e.value)
    assert lib.ndb_b200_quantized_row_bytes(ndb.QUANT_TERNARY, 9) == 3 and lib.ndb_b200_quantized_row_bytes(ndb.QUANT_INT4, 9) == 5
    assert lib.ndb_b200_quantized_row_bytes(ndb.QUANT_FP16, 9) == 18 and lib.ndb_b200_quantized_row_bytes(7, 9) == -1
    if lib.ndb_b200_device_count() > 0:
        return
    cb = np.zeros((2, 4, 2), np.float32)
    for call in (lambda: ndb.cluster_kmeans(X, 2, 2, d[:2]), lambda: ndb.cluster_minibatch_kmeans(X, 2, 4, 2, d),
                 lambda: ndb.pq_train(X, 2, 4, d[:8]), lambda: ndb.pq_encode(X, cb), lambda: ndb.launch_pq_encode(X, cb),
                 lambda: ndb.PqIndex(cb), lambda: ndb.quantize_rows(ndb.QUANT_BINARY, X),
                 lambda: ndb.hamming_knn(np.zeros((4, 2), np.uint8), 16, np.zeros((1, 2), np.uint8), 2)):
        with pytest.raises(ndb.NdbError) as e:
            call()
        assert e.value.code == -3, e.value


def test_product_does_not_import_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "neurondb_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_lib" not in txt and "ndb_oracle" not in txt and "libndb_oracle" not in txt, f
