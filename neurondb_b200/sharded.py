"""Multi-GPU host logic: one process per GPU, torch.distributed for the plumbing.

Which paths shard, and how:

* exact / brute-force kNN -- rows split into contiguous ranges, queries replicated, per-rank
  top-k, all-gather of nq*k (dist, id) pairs, (dist, id) merge (src/util/distributed.c:425-438);
* IVF search -- lists split between the ranks (greedily by member count, or l % world), centroids
  replicated, same all-gather + merge;
* k-means training -- rows split, centroids replicated, one all-reduce of the k*d partial sums and
  k counts (+ one of the cost) per Lloyd iteration;
* quantised scans (PQ asymmetric distance, Hamming) -- encoded rows split into contiguous ranges, codebook / queries
  replicated, per-rank top-k with the local row indices shifted to global ones, the same all-gather + merge;
* HNSW search -- does not shard (graph traversal is global): replicas, queries split, results
  gathered.  HNSW build: one rank builds, the graph is broadcast.

Nothing here computes distances: the per-rank work is passed in as callables (the C-ABI calls on
the GPU box; the oracle in the CPU gloo tests), so the exchange logic is testable without a GPU.
"""
import numpy as np
import torch
import torch.distributed as dist


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def row_range(n, rank, nranks):
    """Contiguous row range [lo, hi) of `rank`: rows [r*n/G, (r+1)*n/G)."""
    return (rank * n) // nranks, ((rank + 1) * n) // nranks


def query_range(nq, rank, nranks):
    """HNSW replicas: the queries are split the same way."""
    return row_range(nq, rank, nranks)


def list_owners(member_counts, nranks):
    """Owner rank of every IVF list: longest list first, each to the rank with the fewest rows so far
    (ties -> lowest rank).  Deterministic, so every rank computes the same map from the replicated
    centroid pages' memberCount."""
    member_counts = np.asarray(member_counts, np.int64)
    owner = np.zeros(member_counts.shape[0], np.int32)
    load = np.zeros(nranks, np.int64)
    for l in np.argsort(-member_counts, kind="stable"):
        r = int(np.argmin(load))
        owner[l] = r
        load[r] += member_counts[l]
    return owner


def gather_topk(dist_t, ids_t):
    """All-gather the per-rank [nq, k] results -> ([world, nq, k] dist, [world, nq, k] ids)."""
    _, nranks = world()
    if nranks == 1:
        return dist_t.unsqueeze(0), ids_t.unsqueeze(0)
    nq = dist_t.shape[0]
    # (concatenated along dim 0: the one output layout every backend accepts)
    all_d = torch.empty((nranks * nq,) + tuple(dist_t.shape[1:]), dtype=dist_t.dtype, device=dist_t.device)
    all_i = torch.empty((nranks * nq,) + tuple(ids_t.shape[1:]), dtype=ids_t.dtype, device=ids_t.device)
    dist.all_gather_into_tensor(all_d, dist_t.contiguous())
    dist.all_gather_into_tensor(all_i, ids_t.contiguous())
    return all_d.view((nranks, nq) + tuple(dist_t.shape[1:])), all_i.view((nranks, nq) + tuple(ids_t.shape[1:]))


def gather_merge(dist_t, ids_t, merge_fn):
    """Per-rank top-k -> global top-k on every rank.  merge_fn([world,nq,k] dist, ids) -> (dist, ids)
    ordered by (dist, id): ndb_b200_merge_topk_dev on the GPU, the oracle's merge in the CPU tests."""
    all_d, all_i = gather_topk(dist_t, ids_t)
    return merge_fn(all_d, all_i)


def gather_query_slices(dist_t, ids_t, nq):
    """HNSW replicas: rank r answered queries query_range(nq, r, world); reassemble [nq, k] everywhere."""
    rank, nranks = world()
    if nranks == 1:
        return dist_t, ids_t
    k = dist_t.shape[1]
    sizes = [query_range(nq, r, nranks) for r in range(nranks)]
    width = max(hi - lo for lo, hi in sizes)
    pd = torch.full((width, k), float("inf"), dtype=dist_t.dtype, device=dist_t.device)
    pi = torch.full((width, k), -1, dtype=ids_t.dtype, device=ids_t.device)
    pd[:dist_t.shape[0]] = dist_t
    pi[:ids_t.shape[0]] = ids_t
    all_d = torch.empty((nranks * width, k), dtype=dist_t.dtype, device=dist_t.device)
    all_i = torch.empty((nranks * width, k), dtype=ids_t.dtype, device=ids_t.device)
    dist.all_gather_into_tensor(all_d, pd)
    dist.all_gather_into_tensor(all_i, pi)
    all_d, all_i = all_d.view(nranks, width, k), all_i.view(nranks, width, k)
    out_d = torch.cat([all_d[r, :hi - lo] for r, (lo, hi) in enumerate(sizes)])
    out_i = torch.cat([all_i[r, :hi - lo] for r, (lo, hi) in enumerate(sizes)])
    return out_d, out_i


def kmeans_train_sharded(step_fn, cost_fn, C0, max_iter=50, tol=0.001):
    """Lloyd's algorithm (kmeans_run, ivf_am.c:2117-2159) over row-sharded samples.

    C0: [k, d] float32 tensor, the replicated initial centroids (kmeans_init: the first k samples of
    the GLOBAL order -- the caller broadcasts them).  step_fn(C) assigns the local rows and returns
    the local per-cluster (sums [k, d] f32, counts [k] i32); cost_fn(C) returns the local cost as a
    1-element f32 tensor.  Per iteration: all-reduce(sums), all-reduce(counts), centroid = sum /
    count (empty cluster -> zeros, :2207-2211), all-reduce(cost), stop when |prev - cost| < tol.
    The sums are added in rank order of the reduction rather than in sample order, so with more than
    one rank the centroids agree with the single-process result to fp32 rounding (1e-5 relative),
    not bit for bit; with one rank they are identical."""
    _, nranks = world()
    C = C0.clone()
    prev = float(np.finfo(np.float32).max)
    cost = 0.0
    iters = 0
    counts = None
    for it in range(max_iter):
        sums, counts = step_fn(C)
        if nranks > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        nonempty = counts > 0
        C = torch.where(nonempty.unsqueeze(1), sums / counts.clamp(min=1).to(sums.dtype).unsqueeze(1), torch.zeros_like(sums))
        c = cost_fn(C)
        if nranks > 1:
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
        cost = float(c.item())
        iters = it + 1
        if abs(np.float32(prev) - np.float32(cost)) < tol:
            break
        prev = cost
    return C, counts, iters, cost


# ---- GPU bindings of the per-rank work (C ABI, device pointers, torch's current stream) ----------
_SIDE = {}


def _on_side_stream(call):
    """Run call(stream_ptr) on a dedicated non-default CUDA stream ordered after, and before, torch's
    current stream (the C ABI treats a NULL stream as "the library's own", which torch cannot see)."""
    dev = torch.cuda.current_device()
    if dev not in _SIDE:
        _SIDE[dev] = torch.cuda.Stream(device=dev)
    side, cur = _SIDE[dev], torch.cuda.current_stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        call(side.cuda_stream)
    cur.wait_stream(side)


def gpu_kmeans_fns(X_t, k):
    """(step_fn, cost_fn) for kmeans_train_sharded over the local rows X_t ([n, d] float32 CUDA tensor):
    ndb_b200_kmeans_shard_step_dev / ndb_b200_kmeans_shard_cost_dev."""
    from . import _lib as L
    from ._lib import check, ptr
    n, d = X_t.shape
    assign = torch.empty(n, dtype=torch.int32, device=X_t.device)
    lib = L.load()

    def step_fn(C):
        C = C.contiguous()
        sums = torch.empty((k, d), dtype=torch.float32, device=X_t.device)
        counts = torch.empty(k, dtype=torch.int32, device=X_t.device)
        _on_side_stream(lambda st: check(lib.ndb_b200_kmeans_shard_step_dev(
            ptr(X_t.data_ptr()), n, d, k, ptr(C.data_ptr()), ptr(assign.data_ptr()), ptr(sums.data_ptr()),
            ptr(counts.data_ptr()), ptr(st))))
        return sums, counts

    def cost_fn(C):
        C = C.contiguous()
        c = torch.empty(1, dtype=torch.float32, device=X_t.device)
        _on_side_stream(lambda st: check(lib.ndb_b200_kmeans_shard_cost_dev(
            ptr(X_t.data_ptr()), n, d, ptr(C.data_ptr()), ptr(assign.data_ptr()), ptr(c.data_ptr()), ptr(st))))
        return c

    step_fn.assign = assign
    return step_fn, cost_fn


def gpu_merge(all_d, all_i):
    """merge_fn for gather_merge: ndb_b200_merge_topk_dev on [world, nq, k] device tensors."""
    from . import _lib as L
    from ._lib import check, ptr
    nranks, nq, k = all_d.shape
    out_d = torch.empty((nq, k), dtype=torch.float32, device=all_d.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=all_d.device)
    _on_side_stream(lambda st: check(L.load().ndb_b200_merge_topk_dev(
        ptr(all_d.data_ptr()), ptr(all_i.data_ptr()), nranks, nq, k, ptr(out_d.data_ptr()), ptr(out_i.data_ptr()), ptr(st))))
    return out_d, out_i


def gpu_knn_sharded(ds, Q_t, k, metric, arith):
    """Exact kNN over row-sharded datasets: `ds` holds this rank's rows (with their GLOBAL ids), Q_t the
    replicated [nq, d] float32 CUDA queries.  Local top-k -> all-gather -> (dist, id) merge."""
    nq = Q_t.shape[0]
    d_t = torch.empty((nq, k), dtype=torch.float32, device=Q_t.device)
    i_t = torch.empty((nq, k), dtype=torch.int64, device=Q_t.device)
    _on_side_stream(lambda st: ds.knn_dev(Q_t.data_ptr(), nq, k, d_t.data_ptr(), i_t.data_ptr(), metric, arith, st))
    return gather_merge(d_t, i_t, gpu_merge)


def global_rows(rows_t, lo):
    """Local row indices of a rank's top-k -> global ones (rows [lo, hi) live on this rank); -1 (past the end) stays."""
    return torch.where(rows_t >= 0, rows_t + lo, rows_t)


def gpu_pq_search_sharded(pq, Q_t, k, lo):
    """ORDER BY pq_asymmetric_distance LIMIT k over row-sharded codes: `pq` holds the codes of rows [lo, hi) of the table,
    Q_t the replicated [nq, dim] float32 CUDA queries.  Local scan -> global rows -> all-gather -> (dist, row) merge."""
    from . import _lib as L
    from ._lib import check, ptr
    nq = Q_t.shape[0]
    d_t = torch.empty((nq, k), dtype=torch.float32, device=Q_t.device)
    r_t = torch.empty((nq, k), dtype=torch.int64, device=Q_t.device)
    _on_side_stream(lambda st: check(L.load().ndb_b200_pq_search_dev(pq.h, ptr(Q_t.data_ptr()), nq, k, ptr(d_t.data_ptr()),
                                                                     ptr(r_t.data_ptr()), ptr(st))))
    return gather_merge(d_t, global_rows(r_t, lo), gpu_merge)


def gpu_hnsw_replicas(h, Q_t, ef, k, mode):
    """HNSW does not shard: every rank holds the whole graph and answers its slice of the queries."""
    rank, nranks = world()
    nq = Q_t.shape[0]
    lo, hi = query_range(nq, rank, nranks)
    d_t = torch.empty((hi - lo, k), dtype=torch.float32, device=Q_t.device)
    i_t = torch.empty((hi - lo, k), dtype=torch.int64, device=Q_t.device)
    if hi > lo:
        q = Q_t[lo:hi].contiguous()
        _on_side_stream(lambda st: h.search_dev(q.data_ptr(), hi - lo, d_t.data_ptr(), i_t.data_ptr(), ef, k, None, mode, st))
    return gather_query_slices(d_t, i_t, nq)
