"""SURVEY 8f-3 / 8f-4 entry points on the GPU with the CPU oracle (the reference's loops, restated and pinned) timed on a
bounded sample next to them: cluster_kmeans, knn_classify, product quantisation (train, encode, asymmetric-distance scan).

All GPU figures are end to end through the C ABI with host buffers (wall clock around the call, copies included); the
PQ scan is additionally timed on the device through ndb_b200_pq_search_dev.  One JSON object per line."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import neurondb_b200 as ndb
import oracle_lib as O
import workloads as W

small = len(sys.argv) > 1 and sys.argv[1] == "small"
ndb.init(0)
cores = os.cpu_count()


def wall(f, reps=1):
    best = None
    for _ in range(reps):
        t = time.perf_counter(); r = f(); dt = time.perf_counter() - t
        best = dt if best is None or dt < best else best
    return best, r


def emit(**kw):
    print(json.dumps(kw), flush=True)


# ---- cluster_kmeans ---------------------------------------------------------------------------------------------
n, dim, k, iters = (20000, 32, 16, 5) if small else (200_000, 64, 64, 5)
X = W.mixture(n, dim, k, 1)
draws = np.random.default_rng(1).integers(0, O.RAND_MAX, k, dtype=np.int64).astype(np.int32)
ndb.cluster_kmeans(X[:2000], 4, 1, draws[:4])                     # warm-up (context, module load)
ndb.cluster_kmeans(X[:5000], k, 2, draws)
tg, (labels, centers, seeds, it) = wall(lambda: ndb.cluster_kmeans(X, k, iters, draws), 2)
walked = ndb.last_kernel_stats()[2]
os.environ["NDB_CKM_SEQUENTIAL"] = "1"
tg_lit, _ = wall(lambda: ndb.cluster_kmeans(X, k, iters, draws))
del os.environ["NDB_CKM_SEQUENTIAL"]
nc = n // 10
tc, (cl, cc, cs, cit) = wall(lambda: O.cluster_kmeans(X[:nc], k, iters, draws))
emit(what="cluster_kmeans", n=n, dim=dim, k=k, lloyd_iterations=it, gpu_s=tg, gpu_rows_per_s=n / tg,
     seeds_that_needed_the_literal_walk=walked, gpu_s_literal_walk_for_every_seed=tg_lit, cpu_sample_rows=nc, cpu_s=tc, cpu_iterations=cit, cpu_s_scaled_to_n=tc * n / nc * (it / max(cit, 1)), cpu_threads=1,
     note="end to end through ndb_b200_cluster_kmeans (host buffers); CPU = orc_cluster_kmeans (the reference's loops) on n/10 rows, scaled linearly")

# ---- cluster_minibatch_kmeans ------------------------------------------------------------------------------------
mb_k, mb_batch, mb_iters = (16, 100, 20) if small else (64, 100, 100)
mdraws = np.random.default_rng(9).integers(0, O.RAND_MAX, mb_k + mb_batch * mb_iters, dtype=np.int64).astype(np.int32)
ndb.cluster_minibatch_kmeans(X[:5000], mb_k, mb_batch, 3, mdraws)
tg, (ml, mc, mused) = wall(lambda: ndb.cluster_minibatch_kmeans(X, mb_k, mb_batch, mb_iters, mdraws), 2)
tc, (cl2, cc2, cused) = wall(lambda: O.cluster_minibatch_kmeans(X[:nc], mb_k, mb_batch, mb_iters, mdraws))
emit(what="cluster_minibatch_kmeans", n=n, dim=dim, k=mb_k, batch=mb_batch, iters=mb_iters, gpu_s=tg, rand_calls=mused,
     cpu_sample_rows=nc, cpu_s=tc, cpu_threads=1,
     note="CPU = orc_cluster_minibatch_kmeans on n/10 rows: its seeding and final assignment scale with n (x10), its %d mini-batch steps do not" % mb_iters)

# ---- knn_classify ------------------------------------------------------------------------------------------------
n, dim, nq, k = (20000, 32, 200, 5) if small else (1_000_000, 64, 2000, 10)
X = W.mixture(n, dim, 64, 2)
Q = W.mixture(nq, dim, 64, 3, centers_seed=2)
lab = np.random.default_rng(2).integers(0, 2, n).astype(np.float64)
ds = ndb.Dataset(dim)
ds.append(X)
ds.knn_classify(lab, Q[:8], k)
tg, cls = wall(lambda: ds.knn_classify(lab, Q, k), 3)
nqc = 4
tc, (ccls, cmean, crow) = wall(lambda: O.knn_ml(X, lab, Q[:nqc], k))
emit(what="knn_classify", n=n, dim=dim, nq=nq, k=k, gpu_s=tg, gpu_qps=nq / tg, cpu_queries=nqc, cpu_s=tc, cpu_qps=nqc / tc, cpu_threads=1,
     same_class_on_sample=bool(np.array_equal(cls[:nqc], ccls)),
     note="rows resident (ndb_b200_dataset), queries and labels from the host per call; CPU = the SQL function's loop + qsort per query")
del ds

# ---- product quantisation -----------------------------------------------------------------------------------------
n, dim, m, ksub, nq, k = (50000, 32, 8, 256, 100, 10) if small else (1_000_000, 128, 16, 256, 1000, 10)
X = W.mixture(n, dim, 256, 4)
Q = W.mixture(nq, dim, 256, 5, centers_seed=4)
ntrain = min(n, 20000)
draws = np.random.default_rng(3).integers(0, O.RAND_MAX, m * ksub, dtype=np.int64).astype(np.int32)
ndb.pq_train(X[:2000], m, ksub, draws, 2)                         # warm-up (kernel load)
tt, cb = wall(lambda: ndb.pq_train(X[:ntrain], m, ksub, draws, 10), 2)
ntc = ntrain // 10
ttc, cbc = wall(lambda: O.pq_train(X[:ntc], m, ksub, draws, 10))
emit(what="pq_train", rows=ntrain, dim=dim, m=m, ksub=ksub, max_iters=10, gpu_s=tt, cpu_sample_rows=ntc, cpu_s=ttc, cpu_s_scaled=ttc * ntrain / ntc,
     cpu_threads=1)
pq = ndb.PqIndex(cb)
ndb.PqIndex(cb).add(X[:2000])
te, codes = wall(lambda: pq.add(X, want_codes=True))
nec = 2000
tec, cc = wall(lambda: O.pq_encode(X[:nec], cb))
emit(what="pq_encode", rows=n, gpu_s=te, gpu_rows_per_s=n / te, cpu_sample_rows=nec, cpu_s=tec, cpu_rows_per_s=nec / tec, cpu_threads=1,
     same_codes_on_sample=bool(np.array_equal(codes[:nec], cc)),
     fp64_ops=3.0 * n * ksub * dim, note="3 rounded fp64 operations per (row, codeword, dimension): fp64-issue bound")
pq.search(Q[:8], k)
ts, (d, r) = wall(lambda: pq.search(Q, k), 3)
# device-timed: queries resident, results left on the device
import torch
Qd = torch.from_numpy(Q).cuda()
dd = torch.empty((nq, k), dtype=torch.float32, device="cuda")
rd = torch.empty((nq, k), dtype=torch.int64, device="cuda")
lib = ndb._lib.load()
st = torch.cuda.Stream()                                          # (a null stream handle selects the library's own stream)
stream = st.cuda_stream
torch.cuda.synchronize()
for _ in range(2):
    assert lib.ndb_b200_pq_search_dev(pq.h, C.c_void_p(Qd.data_ptr()), nq, k, C.c_void_p(dd.data_ptr()), C.c_void_p(rd.data_ptr()), C.c_void_p(stream)) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record(st)
reps = 5
for _ in range(reps):
    assert lib.ndb_b200_pq_search_dev(pq.h, C.c_void_p(Qd.data_ptr()), nq, k, C.c_void_p(dd.data_ptr()), C.c_void_p(rd.data_ptr()), C.c_void_p(stream)) == 0
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
nqc, nrc = 4, min(n, 100_000)
tsc, (cd, cr, _) = wall(lambda: O.pq_knn(Q[:nqc], codes[:nrc], cb, k))
emit(what="pq_scan", rows=n, dim=dim, m=m, ksub=ksub, nq=nq, k=k, e2e_s=ts, e2e_qps=nq / ts, device_ms=ms, device_qps=nq / ms * 1e3,
     code_bytes_per_step=float(n) * m * nq, achieved_code_GBps=float(n) * m * nq / ms / 1e6,
     row_evals_per_s=float(n) * nq / ms * 1e3,
     cpu_sample="%d queries x %d rows" % (nqc, nrc), cpu_s=tsc, cpu_row_evals_per_s=nqc * nrc / tsc, cpu_threads=cores,
     same_as_device=bool(np.array_equal(dd.cpu().numpy(), d) and np.array_equal(rd.cpu().numpy(), r)),
     note="algorithmic bytes = m code bytes per (row, query): the codes are re-read from L2/HBM for every query block; "
          "the table lookups (m fp64 shared-memory reads + adds per row) are the issue bound")

# ---- per-vector quantisers and the Hamming scan --------------------------------------------------------------------
n, dim, nq, k = (50000, 64, 100, 10) if small else (1_000_000, 128, 1000, 10)
X = W.gaussian(n, dim, 6)
Qb = W.gaussian(nq, dim, 7)
ndb.quantize_rows(ndb.QUANT_INT8, X[:1000])
for kind, name in ((ndb.QUANT_INT8, "int8"), (ndb.QUANT_FP16, "fp16"), (ndb.QUANT_BINARY, "binary"), (ndb.QUANT_INT4, "int4")):
    tq, out = wall(lambda: ndb.quantize_rows(kind, X), 2)
    ncq = 20000
    tqc, outc = wall(lambda: O.quantize_rows(kind, X[:ncq]))
    emit(what="quantize_rows", kind=name, rows=n, dim=dim, gpu_s=tq, gpu_rows_per_s=n / tq, e2e_GBps=(X.nbytes + out.nbytes) / tq / 1e9,
         cpu_sample_rows=ncq, cpu_rows_per_s=ncq / tqc, cpu_threads=1, same_bytes_on_sample=bool(np.array_equal(out[:ncq], outc)),
         note="end to end with host buffers: the 4 * dim bytes per row cross PCIe, which bounds it")
bits, qbits = ndb.quantize_rows(ndb.QUANT_BINARY, X), ndb.quantize_rows(ndb.QUANT_BINARY, Qb)
ndb.hamming_knn(bits[:1000], dim, qbits[:8], k)
th, (hd, hi) = wall(lambda: ndb.hamming_knn(bits, dim, qbits, k), 3)
nqc = 8
thc, (cd, ci) = wall(lambda: O.hamming_knn(bits, dim, qbits[:nqc], k))
emit(what="hamming_knn", rows=n, nbits=dim, nq=nq, k=k, e2e_s=th, e2e_qps=nq / th, row_evals_per_s=float(n) * nq / th,
     cpu_queries=nqc, cpu_s=thc, cpu_qps=nqc / thc, cpu_threads=cores, same_on_sample=bool(np.array_equal(hd[:nqc], cd) and np.array_equal(hi[:nqc], ci)),
     note="rows uploaded per call (16 MB); algorithmic bytes = nbits / 8 per (row, query)")
