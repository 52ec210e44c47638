"""Per-kernel view of ONE rank of an N-way striped C4 index on a single GPU (the rank's local search, without the
exchange): python tools/shard_profile.py [N] -- run it under ncu --metrics gpu__time_duration.sum."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
import neurondb_b200 as ndb

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
w = bench.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "c4"]
ndb.init(0)
X, Q = bench.make_data(w)
ix = ndb.IvfIndex(w["dim"], w["lists"], w["metric"])
ix.ivfbuild(X)
ix.ivfinsert(X[0::N], np.arange(0, w["n"], N, dtype=np.int64))
ix.prepare(ndb.ARITH_TENSOR)
nq = w["nq"]
for rep in range(3):
    t = time.perf_counter()
    d, i = ix.search(Q[:nq], w["nprobe"], w["k"], ndb.IVF_FULL, ndb.ARITH_TENSOR)
    dt = time.perf_counter() - t
st = ix.cert_stats(); st["ms_e2e"] = dt * 1e3; st["rows"] = len(ix)
print(json.dumps(st))
ls = ix.list_sizes()
full = ls * N
print("shard list sizes: nonempty %d, rows %d; full-index lists (est.): <2k: %d lists / %.2f of rows, 2k-16k: %d / %.2f, >16k: %d / %.2f, max %d"
      % ((ls > 0).sum(), ls.sum(), (full < 2000).sum(), full[full < 2000].sum() / full.sum(), ((full >= 2000) & (full < 16000)).sum(),
         full[(full >= 2000) & (full < 16000)].sum() / full.sum(), (full >= 16000).sum(), full[full >= 16000].sum() / full.sum(), full.max()))
tiles_striped = np.ceil(ls / 256).sum()
print("tiles per rank striped: %d; ideal %d" % (tiles_striped, np.ceil(full / 256).sum() / N))
