"""N-GPU check of the sharded paths, run under torchrun on a multi-GPU box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multi_gpu_check.py

Row-sharded exact kNN and list-sharded IVF must equal the single-GPU result bit for bit; row-sharded
k-means must match it to fp32 rounding with identical counts; HNSW replicas must return what one
replica returns.  Rank 0 prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import neurondb_b200 as ndb  # noqa: E402
from neurondb_b200 import sharded as S  # noqa: E402
import workloads as W  # noqa: E402


def main():
    rank, nranks, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ndb.init(local)
    res = {"world": nranks}
    n, dim, k = 200_000, 64, 10
    X = W.mixture(n, dim, 64, 401)
    Q = W.mixture(500, dim, 64, 402, centers_seed=401)
    Qt = torch.from_numpy(Q).cuda()
    # exact kNN, rows sharded
    lo, hi = S.row_range(n, rank, nranks)
    ds = ndb.Dataset(dim)
    ds.append(X[lo:hi], np.arange(lo, hi, dtype=np.int64))
    d, i = S.gpu_knn_sharded(ds, Qt, k, ndb.L2, ndb.ARITH_OP_F64)
    full = ndb.Dataset(dim)
    full.append(X)
    wd, wi = full.knn(Q, k, ndb.L2, ndb.ARITH_OP_F64)
    torch.cuda.synchronize()
    res["knn_ids_equal"] = bool(np.array_equal(i.cpu().numpy(), wi))
    res["knn_dist_bits_equal"] = bool(np.array_equal(d.cpu().numpy().view(np.uint32), wd.view(np.uint32)))
    # k-means, rows sharded
    kc = 64
    Xs = torch.from_numpy(X[:20000][S.row_range(20000, rank, nranks)[0]:S.row_range(20000, rank, nranks)[1]]).cuda()
    step_fn, cost_fn = S.gpu_kmeans_fns(Xs, kc)
    C, counts, iters, cost = S.kmeans_train_sharded(step_fn, cost_fn, torch.from_numpy(X[:kc].copy()).cuda(), 50, 0.001)
    wC, _, wcounts, witers, wcost = ndb.kmeans_train(X[:20000], kc)
    res["kmeans_iters"] = [iters, int(witers)]
    res["kmeans_counts_equal"] = bool(np.array_equal(counts.cpu().numpy(), wcounts))
    res["kmeans_centroid_max_rel_err"] = float(np.max(np.abs(C.cpu().numpy() - wC) / np.maximum(np.abs(wC), 1e-3)))
    # IVF, lists sharded by member count
    ix_full = ndb.IvfIndex(dim, 64)
    ix_full.ivfbuild(X)
    lists = ix_full.ivfinsert(X)
    owner = S.list_owners(np.bincount(lists, minlength=64), nranks)
    ix = ndb.IvfIndex(dim, 64)
    ix.set_centroids(ix_full.centroids())
    mine = owner[lists] == rank
    ix.ivfinsert(X[mine], np.flatnonzero(mine).astype(np.int64))
    od = torch.empty((Q.shape[0], k), dtype=torch.float32, device="cuda")
    oi = torch.empty((Q.shape[0], k), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ix.search_dev(Qt.data_ptr(), Q.shape[0], od.data_ptr(), oi.data_ptr(), 8, k)
    ndb.check(ndb._lib.load().ndb_b200_stream_synchronize(None))
    md, mi = S.gather_merge(od, oi, S.gpu_merge)
    fd, fi = ix_full.search(Q, 8, k)
    torch.cuda.synchronize()
    res["ivf_ids_equal"] = bool(np.array_equal(mi.cpu().numpy(), fi))
    res["ivf_dist_bits_equal"] = bool(np.array_equal(md.cpu().numpy().view(np.uint32), fd.view(np.uint32)))
    # HNSW replicas
    h = ndb.HnswIndex(dim, 8, 32, 32)
    h.hnswbuild(X[:5000])
    hd, hi_ = S.gpu_hnsw_replicas(h, Qt, 32, k, ndb.HNSW_BESTFIRST)
    torch.cuda.synchronize()      # a handle's scratch belongs to one stream at a time
    sd, si = h.search(Q, 32, k)
    res["hnsw_replicas_equal"] = bool(np.array_equal(hi_.cpu().numpy(), si))
    flags = torch.tensor([int(all(v for kk, v in res.items() if isinstance(v, bool)))], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    res["all_ranks_ok"] = bool(flags.item())
    if rank == 0:
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()
    ndb.shutdown()


if __name__ == "__main__":
    main()
