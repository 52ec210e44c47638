"""Generate the golden vectors under tests/golden/ (run in the build container, where
/root/reference exists and oracle/_ref/*.so have been built by `make -C oracle`).

operator_distances.npz : outputs of the REFERENCE's own vector_l2_distance / vector_cosine_distance /
    vector_inner_product (NeuronDB/src/vector/vector_distance.c + vector_distance_simd.c compiled
    unmodified under oracle/pgshim), default scalar build and -mavx2 -mfma build, on seeded inputs.
index_paths.npz : outputs of the oracle restatement for the IVF / HNSW / k-means paths (the reference's
    tests pin nothing there -- SURVEY.md 8c -- so these are regression vectors, not reference outputs).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402
import workloads as W  # noqa: E402

DIMS = [1, 2, 3, 4, 7, 8, 9, 15, 16, 17, 31, 33, 96, 128, 768]


def operator_inputs(dim):
    rng = np.random.default_rng(1000 + dim)
    A = rng.standard_normal((48, dim)).astype(np.float32)
    B = rng.standard_normal((48, dim)).astype(np.float32)
    A[0] = 0.0                                   # zero vector (cosine -> 1.0)
    B[1] = A[1]                                  # identical
    A[2] *= 1e18; B[2] *= 1e-18                  # wide dynamic range
    B[3] = -A[3]                                 # opposite (cosine -> 2)
    return A, B


def main():
    assert O.ref_lib() is not None, "build oracle/_ref first (make -C oracle)"
    out = {}
    for dim in DIMS:
        A, B = operator_inputs(dim)
        for metric in (1, 2, 3):
            out["scalar_m%d_d%d" % (metric, dim)] = O.ref_distance_pairs(metric, A, B)
            out["avx2_m%d_d%d" % (metric, dim)] = O.ref_distance_pairs(metric, A, B, avx2=True)
    np.savez_compressed(os.path.join(HERE, "operator_distances.npz"), **out)

    idx = {}
    X = W.mixture(3000, 16, 12, 5)
    Q = W.mixture(40, 16, 12, 6, centers_seed=5)
    C, assign, counts, iters, cost = O.kmeans_train(X[:1200], 12)
    idx.update(km_C=C, km_assign=assign, km_counts=counts, km_iters=np.int32(iters), km_cost=np.float32(cost))
    lists = O.ivf_assign(X, C)
    off, rows = O.lists_from_assignment(lists, 12)
    for lit in (0, 1):
        for metric in (1, 2, 3):
            d, i, cnt = O.ivf_search(X, C, off, rows, Q, 4, 10, strategy=metric, literal=bool(lit))
            idx["ivf_d_l%d_m%d" % (lit, metric)] = d
            idx["ivf_i_l%d_m%d" % (lit, metric)] = i
    idx["ivf_lists"] = lists
    levels = O.hnsw_levels(1500, seed=9)
    idx["hnsw_levels"] = levels
    for mode in (0, 1, 3):            # 3 = the diversity-heuristic extension (NDB_HNSW_SELECT_HEURISTIC)
        g = O.Hnsw(16, 6, 24, 24, capacity=1500)
        g.build(X[:1500], levels, mode)
        e = g.export()
        idx["hnsw_nbr0_b%d" % mode] = e["nbr0"]
        idx["hnsw_cnt_b%d" % mode] = e["cnt"]
        for smode in (0, 1):
            d, n, c = g.search(Q, 24, 10, 1, smode)
            idx["hnsw_d_b%d_s%d" % (mode, smode)] = d
            idx["hnsw_n_b%d_s%d" % (mode, smode)] = n
    d, i = O.knn_exact(X, Q, 10, 1, O.ARITH_OP_F64)
    idx.update(knn_d=d, knn_i=i)
    np.savez_compressed(os.path.join(HERE, "index_paths.npz"), **idx)

    # key extraction: the reference's own fp16_to_float (oracle/_ref/libndb_ref_fp16.so, cut out of
    # src/types/quantization.c by oracle/Makefile) on every binary16 pattern, as float32 bit patterns
    ref16 = O.ref_fp16_lib()
    assert ref16 is not None, "build oracle/_ref first (make -C oracle)"
    import ctypes as C
    ref16.fp16_to_float.restype = C.c_uint32          # read the float's bits back untouched (keeps NaN payloads)
    table = np.array([ref16.fp16_to_float(h) for h in range(65536)], np.uint32)
    ref16.fp16_to_float.restype = C.c_float
    np.savez_compressed(os.path.join(HERE, "fp16_table.npz"), bits=table)
    # the index access methods' leaf functions, from the reference's own code (oracle/extract_ref_leafs.py)
    leafs = O.ref_leafs_lib()
    assert leafs is not None, "build oracle/_ref first (make -C oracle)"
    import test_oracle as T
    ivf, hn = [], []
    for a, b in T._leaf_inputs():
        dim = a.shape[1]
        for s_ in (1, 2):
            ivf.append(np.array([leafs.ref_ivf_distance(a[i], b[i], dim, s_) for i in range(a.shape[0])], np.float32))
        for s_ in (1, 2, 3):
            hn.append(np.array([leafs.ref_hnsw_distance(a[i], b[i], dim, s_) for i in range(a.shape[0])], np.float32))
    Xk = W.mixture(1500, 16, 10, 31)
    kC, ka, kc = O.ref_kmeans_train(Xk, 24)
    np.savez_compressed(os.path.join(HERE, "index_leafs.npz"), ivf_bits=np.concatenate(ivf).view(np.uint32),
                        hnsw_bits=np.concatenate(hn).view(np.uint32), km_C_bits=kC.view(np.uint32), km_assign=ka, km_counts=kc,
                        page_layout=O.ref_page_layout())
    ml_paths()
    print("wrote", sorted(os.listdir(HERE)))


def ml_paths():
    """ml_paths.npz: cluster_kmeans and knn_classify / knn_regress outputs of the REFERENCE's own functions
    (kmeanspp_init, the Lloyd loop of cluster_kmeans, euclidean_distance / compare_samples -- oracle/_ref/
    libndb_ref_leafs.so), with the rand() draws that seeded them.  `python make_golden.py ml` writes only this file."""
    import test_oracle as T
    assert O.ref_leafs_lib() is not None, "build oracle/_ref first (make -C oracle)"
    out = {}
    for n, dim, k, seed in T._ml_cases():
        tag = "n%d" % n
        X = W.mixture(n, dim, max(2, k // 2), seed)
        out["draws_" + tag] = O.libc_rand_draws(seed, k)
        labels, centers, it = O.ref_cluster_kmeans(X, k, 0, seed)
        out["km_labels_" + tag], out["km_center_bits_" + tag], out["km_iters_" + tag] = labels, centers.view(np.uint32), np.int32(it)
        rng = np.random.default_rng(seed)
        Q = W.mixture(30, dim, 5, seed + 1, centers_seed=seed)
        lab = rng.integers(0, 3, n).astype(np.float64)
        cls, mean, _ = O.ref_knn_ml(X, lab, Q, min(k, n))
        out["knn_cls_" + tag], out["knn_mean_" + tag] = cls, mean
    # product quantisation: the reference's train_subspace_kmeans and the loops of pq_encode_vector / pq_asymmetric_distance
    for n, dim, m, ksub, seed in T._pq_cases():
        tag = "pq%d" % n
        X = W.mixture(n, dim, 6, seed)
        Q = W.mixture(12, dim, 6, seed + 1, centers_seed=seed)
        out["draws_" + tag] = O.libc_rand_draws(seed, m * ksub)
        cb = O.ref_pq_train(X, m, ksub, seed, 5)
        codes = O.ref_pq_encode(X, cb)
        out["cb_bits_" + tag], out["codes_" + tag] = cb.view(np.uint32), codes
        out["adc_bits_" + tag] = O.ref_pq_distances(Q, codes, cb).view(np.uint32)
    # cluster_minibatch_kmeans: the reference's own seeding function and main-loop text
    for n, dim, k, batch, iters, seed in T._minibatch_cases():
        X = T._minibatch_rows(n, dim, k, seed)
        out["mb_draws_n%d" % n] = O.libc_rand_draws(seed, k + batch * iters)
        labels, centers = O.ref_cluster_minibatch_kmeans(X, k, batch, iters, seed)
        out["mb_labels_n%d" % n], out["mb_center_bits_n%d" % n] = labels, centers.view(np.uint32)
    # per-vector quantisers: the reference's own quantize_vector_* functions
    for X in T._quant_inputs():
        for kind in (1, 2, 3, 4, 5, 6):
            out["quant_k%d_d%d" % (kind, X.shape[1])] = O.ref_quantize_rows(kind, X)
    np.savez_compressed(os.path.join(HERE, "ml_paths.npz"), **out)


if __name__ == "__main__":
    if sys.argv[1:] == ["ml"]:
        ml_paths()
    else:
        main()
