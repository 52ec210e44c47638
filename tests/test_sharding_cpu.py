"""CPU, world_size 2 over gloo: the N > 1 host logic (lists sharded l % world, per-rank top-k,
all-gather, (dist, id) merge) returns exactly the unsharded result.  The per-rank search runs on
the oracle here; on the GPU box bench.py runs the same exchange over NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
import workloads as W


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X = W.mixture(6000, 16, 24, 11)
    Q = W.mixture(64, 16, 24, 12, centers_seed=11)
    C, _, _, _, _ = O.kmeans_train(X[:2400], 24)
    lists = O.ivf_assign(X, C, nthreads=1)
    # this rank keeps only the lists it owns; the centroids stay replicated
    mine = lists % world == rank
    local_rows = np.flatnonzero(mine)
    off, rows = O.lists_from_assignment(lists[mine], 24)
    d, i, _ = O.ivf_search(X[mine], C, off, rows, Q, 6, 10, ids=local_rows.astype(np.int64), nthreads=1)
    td, ti = torch.from_numpy(d), torch.from_numpy(i)
    gd = [torch.empty_like(td) for _ in range(world)]
    gi = [torch.empty_like(ti) for _ in range(world)]
    dist.all_gather(gd, td)
    dist.all_gather(gi, ti)
    md, mi = O.merge_topk(torch.stack(gd).numpy(), torch.stack(gi).numpy())
    if rank == 0:
        off_all, rows_all = O.lists_from_assignment(lists, 24)
        wd, wi, _ = O.ivf_search(X, C, off_all, rows_all, Q, 6, 10, nthreads=1)
        out.put((np.array_equal(mi, wi), np.array_equal(md.view(np.uint32), wd.view(np.uint32))))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_ivf_merge_equals_unsharded():
    O.lib()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    ids_ok, dist_ok = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ids_ok and dist_ok
