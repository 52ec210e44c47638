"""BASELINE config 5 shape (rows sharded over the GPUs, bf16 tensor-core brute force):
   torchrun --nproc-per-node N tools/c5_bench.py [rows_per_gpu] [nq]
Every rank holds rows_per_gpu x 128 rows (generated from seed + rank, as SURVEY 8d prescribes), the
queries are replicated, each rank answers against its shard with NDB_ARITH_TENSOR, the per-rank top-k
are all-gathered and merged on the device.  Rank 0 prints one JSON line (device time, max over ranks)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neurondb_b200 as ndb  # noqa: E402
from neurondb_b200 import sharded as S  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ndb.init(local)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6_250_000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
    dim, k = 128, 10
    rng = np.random.default_rng(5 + rank)
    ds = ndb.Dataset(dim)
    for s in range(0, n, 1_000_000):
        m = min(1_000_000, n - s)
        ds.append(rng.standard_normal((m, dim), dtype=np.float32), np.arange(rank * n + s, rank * n + s + m, dtype=np.int64))
    Q = torch.from_numpy(np.random.default_rng(99).standard_normal((nq, dim), dtype=np.float32)).cuda()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    for _ in range(3):
        d, i = S.gpu_knn_sharded(ds, Q, k, ndb.L2, ndb.ARITH_TENSOR)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    steps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        d, i = S.gpu_knn_sharded(ds, Q, k, ndb.L2, ndb.ARITH_TENSOR)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    if rank == 0:
        rows = n * world
        print(json.dumps({"workload": "C5 shape: %d x %d rows sharded over %d GPU(s), %d queries, k=%d, bf16 tcgen05" % (rows, dim, world, nq, k),
                          "ms_per_batch": ms, "qps": nq / ms * 1e3, "tflops_aggregate": 2.0 * rows * nq * dim / ms / 1e9,
                          "ids_in_range": bool(((i >= 0) & (i < rows)).all().item())}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
