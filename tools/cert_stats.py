"""Certified-selection statistics of one NDB_ARITH_TENSOR search batch on a bench workload:
   python tools/cert_stats.py [c2|c4|smoke]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
import neurondb_b200 as ndb

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = bench.WORKLOADS[name]
ndb.init(0)
X, Q = bench.make_data(w)
ix = ndb.IvfIndex(w["dim"], w["lists"], w["metric"])
ix.ivfbuild(X); ix.ivfinsert(X); ix.prepare(ndb.ARITH_TENSOR)
nq = w["nq"]
for rep in range(2):
    t = time.perf_counter()
    d, i = ix.search(Q[:nq], w["nprobe"], w["k"], ndb.IVF_FULL, ndb.ARITH_TENSOR)
    dt = time.perf_counter() - t
st = ix.cert_stats()
st["ms_e2e"] = dt * 1e3
st["workload"] = name
st["exact_evals_per_query"] = st["list_exact_evals"] / nq
st["coarse_exact_evals_per_query"] = st["coarse_exact_evals"] / nq
print(json.dumps(st))
