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


def _worker_striped(rank, world, port, out):
    """bench.py's default N > 1 layout: every inverted list striped over the ranks (row i on rank i % world), one
    all-gather of PACKED records -- per rank a [nq*k] f32 distance block padded to 16 bytes, then a [nq*k] int64 id
    block, exactly what ndb_b200_ivf_search_sharded_dev exchanges (csrc/comm.cu RecordLayout) -- and the (dist, id)
    merge."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, dim, L, nq, k = 5000, 12, 16, 37, 10
    X = W.mixture(n, dim, L, 21)
    Q = W.mixture(nq, dim, L, 22, centers_seed=21)
    C, _, _, _, _ = O.kmeans_train(X[:1600], L)
    mine = np.arange(rank, n, world)
    lists = O.ivf_assign(X[mine], C, nthreads=1)              # each rank assigns only its own rows
    off, rows = O.lists_from_assignment(lists, L)
    d, i, _ = O.ivf_search(X[mine], C, off, rows, Q, 5, k, strategy=3, ids=mine.astype(np.int64), nthreads=1)
    ids_off = (nq * k * 4 + 15) // 16 * 16
    stride = (ids_off + nq * k * 8 + 15) // 16 * 16
    rec = np.zeros(stride, np.uint8)
    rec[:nq * k * 4] = d.reshape(-1).view(np.uint8)
    rec[ids_off:ids_off + nq * k * 8] = i.reshape(-1).view(np.uint8)
    allrec = [torch.empty(stride, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(allrec, torch.from_numpy(rec))             # ONE collective
    gd = np.stack([r.numpy()[:nq * k * 4].view(np.float32).reshape(nq, k) for r in allrec])
    gi = np.stack([r.numpy()[ids_off:ids_off + nq * k * 8].view(np.int64).reshape(nq, k) for r in allrec])
    md, mi = O.merge_topk(gd, gi)
    if rank == 0:
        la = O.ivf_assign(X, C, nthreads=1)
        off_all, rows_all = O.lists_from_assignment(la, L)
        wd, wi, _ = O.ivf_search(X, C, off_all, rows_all, Q, 5, k, strategy=3, nthreads=1)
        out.put((np.array_equal(mi, wi), np.array_equal(md.view(np.uint32), wd.view(np.uint32)),
                 np.array_equal(la[mine], lists)))
    dist.barrier()
    dist.destroy_process_group()


def test_striped_lists_packed_record_exchange_equals_unsharded():
    O.lib()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_striped, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    ids_ok, dist_ok, assign_ok = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ids_ok and dist_ok and assign_ok


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


# ---- neurondb_b200.sharded: the exchange logic of the other paths that shard -----------
def _np_step_fns(X, k):
    """Per-rank k-means work on the CPU (numpy) with the same contract as gpu_kmeans_fns."""
    state = {}

    def step_fn(C):
        Cn = C.numpy()
        d2 = ((X[:, None, :] - Cn[None, :, :]) ** 2).sum(-1)
        a = d2.argmin(1)                                   # strict <, lowest index wins ties
        state["a"] = a
        sums = np.zeros_like(Cn)
        np.add.at(sums, a, X)
        return torch.from_numpy(sums), torch.from_numpy(np.bincount(a, minlength=k).astype(np.int32))

    def cost_fn(C):
        Cn = C.numpy()
        return torch.tensor([((X - Cn[state["a"]]) ** 2).sum()], dtype=torch.float32)

    return step_fn, cost_fn


def _worker_sharded(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from neurondb_b200 import sharded as S
    ok = {}
    # k-means over row shards == k-means over all rows (to fp32 rounding), counts exact
    X = W.mixture(3000, 8, 12, 31)
    k = 12
    C0 = torch.from_numpy(X[:k].copy())
    lo, hi = S.row_range(X.shape[0], rank, world)
    step_fn, cost_fn = _np_step_fns(X[lo:hi], k)
    C, counts, iters, cost = S.kmeans_train_sharded(step_fn, cost_fn, C0, 50, 0.001)
    # exact kNN over row shards: per-rank top-k with global ids -> gather -> merge
    Q = W.mixture(40, 8, 12, 32, centers_seed=31)
    d, i = O.knn_exact(X[lo:hi], Q, 10, O.L2, O.ARITH_OP_F64, ids=np.arange(lo, hi, dtype=np.int64), nthreads=1)
    md, mi = S.gather_merge(torch.from_numpy(d), torch.from_numpy(i),
                            lambda ad, ai: tuple(torch.from_numpy(a) for a in O.merge_topk(ad.numpy(), ai.numpy())))
    # quantised scans over row shards (PQ codes, binary rows): local top-k with local rows -> global rows -> gather -> merge
    draws = np.random.default_rng(5).integers(0, O.RAND_MAX, 2 * 16, dtype=np.int64).astype(np.int32)
    cb = O.pq_train(X[:500], 2, 16, draws, 3)
    codes = O.pq_encode(X, cb)
    pd_, pr_, _ = O.pq_knn(Q, codes[lo:hi], cb, 10)
    pmd, pmi = S.gather_merge(torch.from_numpy(pd_), S.global_rows(torch.from_numpy(pr_), lo),
                              lambda ad, ai: tuple(torch.from_numpy(a) for a in O.merge_topk(ad.numpy(), ai.numpy())))
    bits, qbits = O.quantize_rows(O.Q_BINARY, X), O.quantize_rows(O.Q_BINARY, Q)
    hd_, hi_ = O.hamming_knn(bits[lo:hi], 8, qbits, 10)
    hmd, hmi = S.gather_merge(torch.from_numpy(hd_.astype(np.float32)), S.global_rows(torch.from_numpy(hi_), lo),
                              lambda ad, ai: tuple(torch.from_numpy(a) for a in O.merge_topk(ad.numpy(), ai.numpy())))
    # HNSW replicas: each rank answers a slice of the queries, everyone ends up with all of them
    qlo, qhi = S.query_range(Q.shape[0], rank, world)
    fake_d = torch.arange(qlo, qhi, dtype=torch.float32).unsqueeze(1).repeat(1, 3)
    fake_i = torch.arange(qlo, qhi, dtype=torch.int64).unsqueeze(1).repeat(1, 3)
    gd, gi = S.gather_query_slices(fake_d, fake_i, Q.shape[0])
    if rank == 0:
        s1, c1 = _np_step_fns(X, k)
        # single-process run of the same driver (world() is 2 here, so drive it by hand)
        Cs = C0.clone(); prev = np.finfo(np.float32).max
        for it in range(50):
            sums, cnt = s1(Cs)
            Cs = torch.where((cnt > 0).unsqueeze(1), sums / cnt.clamp(min=1).float().unsqueeze(1), torch.zeros_like(sums))
            c = float(c1(Cs).item())
            if abs(np.float32(prev) - np.float32(c)) < 0.001:
                break
            prev = c
        ok["kmeans_iters"] = iters == it + 1
        ok["kmeans_counts"] = bool(torch.equal(counts, cnt))
        ok["kmeans_centroids"] = bool(torch.allclose(C, Cs, rtol=1e-5, atol=1e-6))
        wd, wi = O.knn_exact(X, Q, 10, O.L2, O.ARITH_OP_F64, nthreads=1)
        ok["knn_ids"] = np.array_equal(mi.numpy(), wi)
        ok["knn_dist"] = np.array_equal(md.numpy().view(np.uint32), wd.view(np.uint32))
        wpd, wpr, _ = O.pq_knn(Q, codes, cb, 10)
        ok["pq_rows"] = np.array_equal(pmi.numpy(), wpr)
        ok["pq_dist"] = np.array_equal(pmd.numpy().view(np.uint32), wpd.view(np.uint32))
        whd, whi = O.hamming_knn(bits, 8, qbits, 10)
        ok["hamming"] = np.array_equal(hmi.numpy(), whi) and np.array_equal(hmd.numpy().astype(np.int32), whd)
        ok["replicas"] = bool(torch.equal(gi[:, 0], torch.arange(Q.shape[0])) and gd.shape == (Q.shape[0], 3))
        out.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_kmeans_knn_and_replicas_world2():
    O.lib()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    ok = out.get(timeout=90)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok.values()), ok


def test_list_owners_balance_and_determinism():
    from neurondb_b200 import sharded as S
    sizes = np.random.default_rng(3).integers(0, 5000, size=257)
    own = S.list_owners(sizes, 4)
    assert np.array_equal(own, S.list_owners(sizes, 4))
    load = np.bincount(own, weights=sizes, minlength=4)
    assert load.max() - load.min() <= sizes.max()
    assert S.row_range(10, 0, 3) == (0, 3) and S.row_range(10, 2, 3) == (6, 10)
