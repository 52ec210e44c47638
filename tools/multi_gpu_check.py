"""N-GPU check of the sharded paths through the library's own communicator
(ndb_b200_comm_*), run under torchrun on a multi-GPU box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multi_gpu_check.py

Row-sharded exact kNN and the sharded IVF searches (every list striped over the ranks, or whole lists
split) must equal the single-GPU result bit for bit; row-sharded k-means must match it to fp32 rounding
with identical counts; a broadcast HNSW graph must answer like the graph it was copied from.
Rank 0 prints one JSON line."""
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
import workloads as W  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def main():
    rank, nranks, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ndb.init(local)
    ndb.comm_init_torch()
    res = {"world": nranks, "comm_nranks": ndb.comm_nranks(), "nccl": int(ndb._lib.load().ndb_b200_comm_nccl_version())}
    n, dim, k = 200_000, 64, 10
    X = W.mixture(n, dim, 64, 401)
    Q = W.mixture(500, dim, 64, 402, centers_seed=401)
    Qt = torch.from_numpy(Q).cuda()
    nq = Q.shape[0]
    od = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    oi = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()

    # exact kNN, rows sharded (contiguous ranges), fp64 operator arithmetic
    lo, hi = (rank * n) // nranks, ((rank + 1) * n) // nranks
    ds = ndb.Dataset(dim)
    ds.append(X[lo:hi], np.arange(lo, hi, dtype=np.int64))
    ds.knn_sharded_dev(Qt.data_ptr(), nq, k, od.data_ptr(), oi.data_ptr(), ndb.L2, ndb.ARITH_OP_F64)
    ndb.check(ndb._lib.load().ndb_b200_stream_synchronize(None))
    full = ndb.Dataset(dim)
    full.append(X)
    wd, wi = full.knn(Q, k, ndb.L2, ndb.ARITH_OP_F64)
    res["knn_ids_equal"] = bool(np.array_equal(oi.cpu().numpy(), wi))
    res["knn_dist_bits_equal"] = bool(np.array_equal(bits(od.cpu().numpy()), bits(wd)))

    # IVF: single-GPU reference, then striped rows and whole lists
    ix_full = ndb.IvfIndex(dim, 64)
    ix_full.ivfbuild(X)
    ix_full.ivfinsert(X)
    fd, fi = ix_full.search(Q, 8, k)
    td, ti = ix_full.search(Q, 8, k, ndb.IVF_FULL, ndb.ARITH_TENSOR)
    for mode in ("rows", "lists"):
        ix = ndb.IvfIndex(dim, 64)
        ix.set_centroids(ix_full.centroids())
        if mode == "rows":
            ix.ivfinsert(X[rank::nranks], np.arange(rank, n, nranks, dtype=np.int64))
        else:
            ix.set_shard(rank, nranks)
            ix.ivfinsert(X, np.arange(n, dtype=np.int64))
        ix.search_sharded_dev(Qt.data_ptr(), nq, od.data_ptr(), oi.data_ptr(), 8, k)
        ndb.check(ndb._lib.load().ndb_b200_stream_synchronize(None))
        res["ivf_%s_ids_equal" % mode] = bool(np.array_equal(oi.cpu().numpy(), fi))
        res["ivf_%s_dist_bits_equal" % mode] = bool(np.array_equal(bits(od.cpu().numpy()), bits(fd)))
        # host-pointer form, tensor arithmetic: the union of the ranks' candidates can only be better
        hd, hi_ = ix.search_sharded(Q, 8, k, ndb.IVF_FULL, ndb.ARITH_TENSOR)
        res["ivf_%s_tensor_ids_vs_fp32" % mode] = float((hi_ == fi).mean())
        res["ivf_%s_tensor_ids_vs_1gpu_tensor" % mode] = float((hi_ == ti).mean())
        same = hi_ == fi
        res["ivf_%s_tensor_dist_bits_equal_where_ids_agree" % mode] = bool(np.array_equal(bits(hd[same]), bits(fd[same])))
        res["ivf_%s_rows_this_rank" % mode] = len(ix)

    # k-means, rows sharded: the whole Lloyd loop inside the library, all-reduce per iteration
    kc, ns = 64, 20000
    slo, shi = (rank * ns) // nranks, ((rank + 1) * ns) // nranks
    Xs = torch.from_numpy(X[slo:shi]).cuda()
    C = torch.from_numpy(X[:kc].copy()).cuda()
    assign = torch.empty(shi - slo, dtype=torch.int32, device="cuda")
    counts = torch.empty(kc, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    iters, cost = ndb.kmeans_train_sharded_dev(Xs.data_ptr(), shi - slo, dim, kc, C.data_ptr(), assign.data_ptr(), counts.data_ptr())
    ndb.check(ndb._lib.load().ndb_b200_stream_synchronize(None))
    wC, _, wcounts, witers, wcost = ndb.kmeans_train(X[:ns], kc)
    res["kmeans_iters"] = [iters, int(witers)]
    res["kmeans_counts_equal"] = bool(np.array_equal(counts.cpu().numpy(), wcounts))
    res["kmeans_centroid_max_rel_err"] = float(np.max(np.abs(C.cpu().numpy() - wC) / np.maximum(np.abs(wC), 1e-3)))
    res["kmeans_cost_rel_err"] = float(abs(cost - wcost) / max(abs(wcost), 1e-9))

    # HNSW: rank 0 builds, the graph is broadcast, every replica answers like the original
    h = ndb.HnswIndex(dim, 8, 32, 32)
    if rank == 0:
        h.hnswbuild(X[:5000])
    h.broadcast(root=0)
    sd, si = h.search(Q, 32, k)
    ref = [sd.copy(), si.copy()]
    dist.broadcast_object_list(ref, src=0)
    res["hnsw_broadcast_equal"] = bool(np.array_equal(si, ref[1]) and np.array_equal(bits(sd), bits(ref[0])) and len(h) == 5000)

    res["exchange"] = "peer memory (IPC windows, NVLink stores)" if ndb._lib.load().ndb_b200_comm_exchange_is_p2p() else "ncclAllGather"
    ok = all(v for kk, v in res.items() if isinstance(v, bool)) and res["comm_nranks"] == nranks
    flags = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    res["all_ranks_ok"] = bool(flags.item())
    if rank == 0:
        print(json.dumps(res))
    ndb.comm_shutdown()
    dist.barrier()
    dist.destroy_process_group()
    ndb.shutdown()


if __name__ == "__main__":
    main()
