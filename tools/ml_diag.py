"""Repeatability of the end-to-end figures of tools/ml_bench.py: each call timed several times in one process."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import neurondb_b200 as ndb
import workloads as W
ndb.init(0)


def reps(name, f, n=5):
    out = []
    for _ in range(n):
        l0 = ndb.launch_count(); t = time.perf_counter(); f(); out.append((time.perf_counter() - t) * 1e3)
    print(name, "ms:", " ".join("%.1f" % x for x in out), "launches/call:", ndb.launch_count() - l0, flush=True)


X = W.mixture(200_000, 64, 64, 1)
draws = np.random.default_rng(1).integers(0, 2147483647, 4096, dtype=np.int64).astype(np.int32)
reps("cluster_kmeans 200k x 64, k 64, 5 it", lambda: ndb.cluster_kmeans(X, 64, 5, draws[:64]))
X2 = W.mixture(20_000, 128, 256, 4)
reps("pq_train 20k x 128, m 16, ksub 256, 10 it", lambda: ndb.pq_train(X2, 16, 256, draws, 10))
X3 = W.gaussian(1_000_000, 128, 6)
reps("quantize int8 1M x 128", lambda: ndb.quantize_rows(ndb.QUANT_INT8, X3), 4)
bits = ndb.quantize_rows(ndb.QUANT_BINARY, X3)
qb = ndb.quantize_rows(ndb.QUANT_BINARY, W.gaussian(1000, 128, 7))
reps("hamming 1M x 128 bit, 1000 q", lambda: ndb.hamming_knn(bits, 128, qb, 10))
