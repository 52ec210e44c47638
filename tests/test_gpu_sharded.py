"""GPU: the per-rank halves of the sharded paths through the C ABI, one rank.
With a single rank every sharded driver must reproduce the unsharded result bit for bit; the N > 1
exchange logic is covered on CPU (test_sharding_cpu.py, gloo) and by tools/multi_gpu_check.py."""
import numpy as np
import pytest
import torch

import workloads as W

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_sharded_kmeans_one_rank_equals_kmeans_train(ndb, orc):
    from neurondb_b200 import sharded as S
    X = W.mixture(5000, 24, 40, 71)
    k = 40
    Xt = torch.from_numpy(X).cuda()
    step_fn, cost_fn = S.gpu_kmeans_fns(Xt, k)
    C, counts, iters, cost = S.kmeans_train_sharded(step_fn, cost_fn, Xt[:k].clone(), 50, 0.001)
    oC, oassign, ocounts, oiters, ocost = orc.kmeans_train(X, k)
    assert iters == oiters
    assert np.array_equal(counts.cpu().numpy(), ocounts)
    assert np.array_equal(bits(C.cpu().numpy()), bits(oC))
    assert np.float32(cost) == np.float32(ocost)
    assert np.array_equal(step_fn.assign.cpu().numpy(), oassign)


def test_sharded_knn_and_replicas_one_rank(ndb, orc):
    from neurondb_b200 import sharded as S
    X = W.gaussian(4000, 32, 5)
    Q = W.gaussian(50, 32, 6)
    ds = ndb.Dataset(32)
    ds.append(X)
    d, i = S.gpu_knn_sharded(ds, torch.from_numpy(Q).cuda(), 10, ndb.L2, ndb.ARITH_OP_F64)
    od, oi = orc.knn_exact(X, Q, 10, orc.L2, orc.ARITH_OP_F64)
    torch.cuda.synchronize()
    assert np.array_equal(i.cpu().numpy(), oi) and np.array_equal(bits(d.cpu().numpy()), bits(od))
    h = ndb.HnswIndex(32, 8, 32, 32)
    h.hnswbuild(X[:1500])
    hd, hi = S.gpu_hnsw_replicas(h, torch.from_numpy(Q).cuda(), 32, 10, ndb.HNSW_BESTFIRST)
    torch.cuda.synchronize()      # a handle's scratch belongs to one stream at a time (_dev calls are asynchronous)
    wd, wi = h.search(Q, 32, 10)
    assert np.array_equal(hi.cpu().numpy(), wi) and np.array_equal(bits(hd.cpu().numpy()), bits(wd))
