"""Quick device timing of the main kernels (development aid; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import neurondb_b200 as ndb
import workloads as W

ndb.init(0)
ndb.set_timing(True)
what = sys.argv[1:] or ["c1", "c2"]
if "c1" in what:
    X = W.gaussian(100_000, 128, 1234); Q = W.gaussian(1000, 128, 4321)
    ds = ndb.Dataset(128); ds.append(X)
    for arith, nm in ((ndb.ARITH_OP_F64, "op_f64"), (ndb.ARITH_IVF_F32, "ivf_f32"), (ndb.ARITH_FAST, "fast")):
        for _ in range(3):
            t = time.time(); ds.knn(Q, 10, 1, arith); e2e = time.time() - t
        ms, b, ev = ndb.last_kernel_stats()
        print(f"C1 exact {nm}: scan kernel {ms:.3f} ms  e2e {e2e*1e3:.2f} ms  evals {ev:.3g}  -> {1000/e2e:.0f} QPS e2e, {ev/ms/1e6:.1f} Gdist/s")
if "c2" in what:
    n = int(os.environ.get("C2_N", 1_000_000))
    X = W.mixture(n, 128, 1024, 2024); Q = W.mixture(10_000, 128, 1024, 2025, centers_seed=2024)
    ix = ndb.IvfIndex(128, 1024)
    t = time.time(); ix.ivfbuild(X); t1 = time.time(); ix.ivfinsert(X); t2 = time.time()
    print(f"C2 build: train {t1-t:.2f}s insert {t2-t1:.2f}s  list sizes min/mean/max {ix.list_sizes().min()}/{ix.list_sizes().mean():.0f}/{ix.list_sizes().max()}")
    for arith, nm in ((ndb.ARITH_IVF_F32, "ivf_f32"), (ndb.ARITH_FAST, "fast")):
        for _ in range(3):
            t = time.time(); d, i = ix.search(Q, 16, 10, arith=arith); e2e = time.time() - t
        ms, b, ev = ndb.last_kernel_stats()
        print(f"C2 ivf {nm}: scan kernel {ms:.3f} ms e2e {e2e*1e3:.2f} ms algo {b/1e9:.2f} GB -> {b/ms/1e6:.0f} GB/s algorithmic, {10000/e2e:.0f} QPS e2e, {10000/ms*1e3:.0f} QPS kernel")
    gt = W.exact_ground_truth(X, Q[:500], 10)
    hit = np.mean([len(set(a) & set(b)) / 10 for a, b in zip(i[:500], gt)])
    print("recall@10 (500 queries):", hit)
