"""Debug helper: tensor IVF search on one GPU holding shard `r` of `w` of the lists."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import neurondb_b200 as ndb  # noqa: E402
import workloads as W  # noqa: E402

r, w = int(sys.argv[1]), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 200_000
lists = int(sys.argv[4]) if len(sys.argv) > 4 else 256
ndb.init(0)
X = W.mixture(n, 128, lists, 2024)
Q = W.mixture(2000, 128, lists, 2025, centers_seed=2024)
ix = ndb.IvfIndex(128, lists)
ix.set_shard(r, w)
ix.ivfbuild(X)
ix.ivfinsert(X)
d0, i0 = ix.search(Q, 16, 10, ndb.IVF_FULL, ndb.ARITH_IVF_F32)
d1, i1 = ix.search(Q, 16, 10, ndb.IVF_FULL, ndb.ARITH_TENSOR)
print("rows kept", len(ix), "ids equal", float((i0 == i1).mean()))
