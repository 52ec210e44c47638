"""Brute-force kNN on the tensor cores (BASELINE config 5 per-GPU shard shape): device time, TFLOP/s."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import neurondb_b200 as ndb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6_250_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
dim, k = (int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 128), 10
ndb.init(0)
ndb.set_timing(True)
rng = np.random.default_rng(5)
ds = ndb.Dataset(dim)
chunk = 1_000_000
for s in range(0, n, chunk):
    ds.append(rng.standard_normal((min(chunk, n - s), dim), dtype=np.float32))
Q = rng.standard_normal((nq, dim), dtype=np.float32)
for arith, nm in ((ndb.ARITH_TENSOR, "tensor bf16"), (ndb.ARITH_FAST, "fp32 FFMA")):
    if arith == ndb.ARITH_FAST and n > 1_000_000 and "--all" not in sys.argv:
        continue
    for _ in range(3):
        t = time.time(); d, i = ds.knn(Q, k, ndb.L2, arith); e2e = time.time() - t
    ms, b, ev = ndb.last_kernel_stats()
    flops = 2.0 * n * nq * dim
    print(f"{nm}: n={n} nq={nq}: kernel {ms:.3f} ms -> {flops/ms/1e9:.1f} TFLOP/s, {nq/ms*1e3:.0f} QPS (e2e {nq/e2e:.0f}), "
          f"stored bytes/launch {b/1e9:.2f} GB -> {b/ms/1e6:.0f} GB/s")
