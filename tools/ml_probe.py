"""One full-size launch of a quantised-path kernel as the FIRST launch of that kernel in the process, for
`ncu -k regex:<kernel> --launch-count 1` (no warm-up launches to skip).  Usage: ml_probe.py pq | ham | ckm"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import neurondb_b200 as ndb

part = sys.argv[1]
ndb.init(0)
rng = np.random.default_rng(1)
if part == "pq":                                   # pq_adc_kernel<1>: 1 M rows, m = 16, ksub = 256, 1000 queries
    cb = rng.standard_normal((16, 256, 8), dtype=np.float32)
    pq = ndb.PqIndex(cb)
    pq.add_codes(rng.integers(0, 256, (1_000_000, 16)).astype(np.int16))
    d, r = pq.search(rng.standard_normal((1000, 128), dtype=np.float32), 10)
    print("pq", d[0, :3], r[0, :3])
elif part == "ham":                                # hamming_topk_kernel<1>: 1 M rows of 128 bits, 1000 queries
    rows = rng.integers(0, 256, (1_000_000, 16)).astype(np.uint8)
    q = rng.integers(0, 256, (1000, 16)).astype(np.uint8)
    d, i = ndb.hamming_knn(rows, 128, q, 10)
    print("ham", d[0, :3], i[0, :3])
elif part == "ckm":                                # ckm_assign_kernel<4>: 200 k rows x 64 dims against 64 centres
    X = rng.standard_normal((200_000, 64), dtype=np.float32)
    draws = rng.integers(0, 2147483647, 64, dtype=np.int64).astype(np.int32)
    labels, centers, seeds, it = ndb.cluster_kmeans(X, 64, 2, draws)
    print("ckm", it, labels[:4])
