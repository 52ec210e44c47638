"""HNSW (BASELINE config 3 shape, scaled): build seconds, search QPS, evaluations per query and
recall@10 on the GPU; the reference-literal CPU build/search (oracle) timed next to it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import neurondb_b200 as ndb
import oracle_lib as O
import workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
n_cpu = int(sys.argv[3]) if len(sys.argv) > 3 else 20_000
m, efc, efs, nq = 16, int(os.environ.get("HNSW_EFC", "64")), 40, 10_000
ndb.init(0)
ndb.set_timing(True)
if os.environ.get("HNSW_DATA") == "mixture":        # clustered data (the C2 mixture): a regime where graph search can reach high recall
    X = W.mixture(n, dim, 1024, 2024)
    Q = W.mixture(nq, dim, 1024, 2025, centers_seed=2024)
else:                                               # structureless unit vectors
    X = W.normalised(n, dim, 768)
    Q = W.normalised(nq, dim, 769)
efs = int(os.environ.get("HNSW_EF", efs))
levels = O.hnsw_levels(n, seed=768)
h = ndb.HnswIndex(dim, m, efc, efs, ndb.COSINE)
select = ndb.HNSW_SELECT_HEURISTIC if os.environ.get("HNSW_SELECT") == "heuristic" else ndb.HNSW_SELECT_CLOSEST
t = time.time(); h.hnswbuild(X, levels=levels, select=select, batch=int(os.environ.get("HNSW_BATCH", "0"))); tb = time.time() - t
print(f"GPU build n={n} dim={dim} M={m} efC={efc} select={'heuristic' if select else 'closest (reference)'}: {tb:.2f} s "
      f"({n/tb:.0f} inserts/s), {h.last_evals()/n:.0f} evals/insert")
gt = W.exact_ground_truth(X, Q[:500], 10)
for extra_ef in [int(e) for e in os.environ.get("HNSW_EFS", "").split(",") if e]:
    for _ in range(2):
        t = time.time(); d, i = h.search(Q, extra_ef, 10, 1, ndb.HNSW_BESTFIRST); e2e = time.time() - t
    ms, _, _ = ndb.last_kernel_stats()
    print(f"GPU search best-first L2 ef={extra_ef}: kernel {ms:.3f} ms -> {nq/ms*1e3:.0f} QPS ({nq/e2e:.0f} e2e), "
          f"{h.last_evals()/nq:.0f} evals/query, recall@10 {O.recall_at_k(i[:500], gt):.4f}")
for mode, nm in ((ndb.HNSW_BESTFIRST, "best-first"), (ndb.HNSW_LITERAL, "literal")):
    for strategy, sn in ((1, "L2"), (2, "cosine")):
        for _ in range(3):
            t = time.time(); d, i = h.search(Q, efs, 10, strategy, mode); e2e = time.time() - t
        ms, _, _ = ndb.last_kernel_stats()
        ev = h.last_evals() / nq
        rec = O.recall_at_k(i[:500], gt)
        bytes_q = ev * (dim * 4 + 2 * m * 4)
        print(f"GPU search {nm:10s} {sn:6s} ef={efs}: kernel {ms:.3f} ms -> {nq/ms*1e3:.0f} QPS ({nq/e2e:.0f} e2e), {ev:.0f} evals/query, "
              f"{bytes_q*nq/ms/1e6:.0f} GB/s gathered, recall@10 {rec:.4f}")
if n_cpu:
    Xc, lc = X[:n_cpu], levels[:n_cpu]
    gtc = W.exact_ground_truth(Xc, Q[:500], 10)
    for bmode, bn in ((0, "reference-literal"), (1, "per-level")):
        g = O.Hnsw(dim, m, efc, efs, capacity=n_cpu, native=True)
        t = time.time(); g.build(Xc, lc, bmode); tcb = time.time() - t
        for smode, sn in ((0, "literal"), (1, "best-first")):
            t = time.time(); d, nn, _ = g.search(Q[:2000], efs, 10, 1, smode, nthreads=os.cpu_count()); ts = time.time() - t
            rec = O.recall_at_k(nn[:500].astype(np.int64), gtc)
            print(f"CPU oracle n={n_cpu} build {bn} {tcb:.1f} s ({n_cpu/tcb:.0f} inserts/s, 1 thread); search {sn}: {2000/ts:.0f} QPS "
                  f"({os.cpu_count()} threads), recall@10 {rec:.4f}")
    hg = ndb.HnswIndex(dim, m, efc, efs)
    t = time.time(); hg.hnswbuild(Xc, levels=lc); tg = time.time() - t
    d, i = hg.search(Q[:2000], efs, 10, 1, ndb.HNSW_BESTFIRST)
    print(f"GPU n={n_cpu}: build {tg:.2f} s; best-first recall@10 {O.recall_at_k(i[:500], gtc):.4f}")
