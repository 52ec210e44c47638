#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native NeuronDB vector-search hot path.

Workload (BASELINE.json configs[1], "C2"): IVFFlat L2, 1M x 128 synthetic vectors (1024-component
Gaussian mixture), lists=1024, nprobe=16, k=10, one 10k-query batch per step.
A step = ivfSelectClusters + ivfCollectCandidates for the whole batch (NeuronDB/src/index/ivf_am.c
:1597-1909) through the C ABI of libndb_b200.so.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c4]

Prints ONE JSON line (see the keys below).  `value` = QPS with the query batch already resident in
HBM (device-pointer entry point, CUDA events, max over ranks); `e2e` = the same metric through the
host-pointer entry point with pinned host buffers, H2D and D2H inside the timed region.
N > 1 (torchrun): inverted lists are sharded l % N across ranks, every rank answers the whole batch
against its shard, per-rank top-k are exchanged with an NCCL all-gather and merged on the device by
(dist, id) -- strong scaling of the same workload.

--impl reference times the reference's CPU algorithm for the same step (the oracle restatement
compiled -O3 -march=native, OpenMP over queries, all host cores) on a bounded sample of queries.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (rows, dim, lists, nprobe, k, nq, metric, mixture components, seed)
    "c2": dict(n=1_000_000, dim=128, lists=1024, nprobe=16, k=10, nq=10_000, metric=1, comps=1024, seed=2024,
               label="C2: IVFFlat L2 1Mx128 lists=1024 nprobe=16 k=10, 10k-query batch"),
    "c4": dict(n=10_000_000, dim=96, lists=4096, nprobe=32, k=10, nq=10_000, metric=3, comps=4096, seed=96,
               label="C4: IVFFlat inner-product 10Mx96 lists=4096 nprobe=32 k=10, 10k-query batch"),
    "smoke": dict(n=50_000, dim=64, lists=64, nprobe=8, k=10, nq=1000, metric=1, comps=64, seed=7,
                  label="smoke: IVFFlat L2 50kx64"),
}


def make_data(w, qstream=0):
    """Rows, and 4 batches of queries from the same mixture; qstream picks a different query stream
    (the replicas of a multi-GPU job each answer their own)."""
    import workloads as W
    X = W.mixture(w["n"], w["dim"], w["comps"], w["seed"])
    Q = W.mixture(w["nq"] * 4, w["dim"], w["comps"], w["seed"] + 1 + 7919 * qstream, centers_seed=w["seed"])
    return X, Q


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # clocks under load = the upper half of the samples (idle samples before/after are low)
        sm.sort()
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback"


def cpu_reference_qps(w, X, Q, Cn, nq_sample, threads, native=True):
    """The reference's CPU algorithm for the same step on `nq_sample` queries (oracle port)."""
    import oracle_lib as O
    assign = O.ivf_assign(X, Cn, nthreads=threads, native=native)
    off, rows = O.lists_from_assignment(assign, w["lists"])
    Qs = np.ascontiguousarray(Q[:nq_sample])
    t = time.perf_counter()
    O.ivf_search(X, Cn, off, rows, Qs, w["nprobe"], w["k"], strategy=w["metric"], literal=False, nthreads=threads,
                 native=native)
    dt = time.perf_counter() - t
    return nq_sample / dt, dt, (off, rows)


def run_reference(args, w):
    """--impl reference: the CPU path on the host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as O
    O.build_oracle()
    X, Q = make_data(w)
    ns = O.lib().orc_ivf_train_samples(X.shape[0], w["lists"])
    t0 = time.perf_counter()
    Cn, _, _, iters, _ = O.kmeans_train(X[:ns], w["lists"])
    train_s = time.perf_counter() - t0
    cores = os.cpu_count() or 1
    # bounded sample: sized for roughly 2-4 s of all-core work per step
    sample = min(w["nq"], 2000)
    assign = O.ivf_assign(X, Cn, nthreads=cores, native=True)
    off, rows = O.lists_from_assignment(assign, w["lists"])
    times = []
    for s in range(args.warmup + args.steps):
        Qs = np.ascontiguousarray(Q[(s % 4) * w["nq"]:(s % 4) * w["nq"] + sample])
        t = time.perf_counter()
        O.ivf_search(X, Cn, off, rows, Qs, w["nprobe"], w["k"], strategy=w["metric"], literal=False, nthreads=cores,
                     native=True)
        if s >= args.warmup:
            times.append(time.perf_counter() - t)
    dt = float(np.mean(times))
    qps = sample / dt
    line = {
        "impl": "reference", "metric": "QPS@recall@10>=0.95", "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["label"], "rows": w["n"], "dim": w["dim"], "lists": w["lists"], "nprobe": w["nprobe"],
                   "k": w["k"], "queries_per_step": sample},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": "%d queries of the 10k batch per step, oracle/ndb_oracle.c -O3 -march=native, "
                                   "OpenMP over queries; excludes PostgreSQL executor/bufmgr overhead" % sample,
                         "kmeans_train_s": train_s, "kmeans_iters": iters},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--arith", default="tensor", choices=["ivf_f32", "fast", "tensor"],
                    help="tensor: tcgen05 bf16 candidate selection + fp32 re-rank (default); ivf_f32: the reference's fp32 "
                         "arithmetic end to end, bit-exact distances and ids")
    ap.add_argument("--shard", default="queries", choices=["queries", "lists"],
                    help="N > 1: 'queries' = index replicated, every GPU answers its own query batches, no data-path "
                         "collective (weak scaling); 'lists' = inverted lists split between the GPUs, queries replicated, "
                         "NCCL all-gather of the per-rank top-k + device merge (strong scaling; the capacity mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    w = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference(args, w)
        return

    import torch
    import neurondb_b200 as ndb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ndb.init(local)
    arith = {"ivf_f32": ndb.ARITH_IVF_F32, "fast": ndb.ARITH_FAST, "tensor": ndb.ARITH_TENSOR}[args.arith]

    gather = world > 1 and args.shard == "lists"        # lists split between the ranks -> results merged
    replicas = world if (world > 1 and not gather) else 1
    X, Q = make_data(w, rank if replicas > 1 else 0)
    nq, k, dim = w["nq"], w["k"], w["dim"]
    ix = ndb.IvfIndex(dim, w["lists"], w["metric"])
    if gather:
        ix.set_shard(rank, world)
    t0 = time.perf_counter()
    ix.ivfbuild(X)                     # k-means on the first min(10000, lists*100) rows (ivf_am.c:580)
    t1 = time.perf_counter()
    ix.ivfinsert(X)                    # list assignment + append for every row (ivf_am.c:797-1167)
    t2 = time.perf_counter()
    build = {"train_s": t1 - t0, "insert_s": t2 - t1}

    # a real (non-NULL) stream: the C ABI treats a NULL stream as "the library's own stream", and
    # torch.cuda.Event only sees work queued on torch's current stream
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    qd = [torch.from_numpy(Q[i * nq:(i + 1) * nq]).cuda() for i in range(4)]
    out_d = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    out_i = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    fin_d = fin_i = None
    if gather:
        all_d = torch.empty((world, nq, k), dtype=torch.float32, device="cuda")
        all_i = torch.empty((world, nq, k), dtype=torch.int64, device="cuda")
        fin_d = torch.empty_like(out_d)
        fin_i = torch.empty_like(out_i)

    def step(s, arith=arith):
        q = qd[0 if os.environ.get("NDB_BENCH_ONE_BATCH") else s % 4]
        ix.search_dev(q.data_ptr(), nq, out_d.data_ptr(), out_i.data_ptr(), w["nprobe"], k, ndb.IVF_FULL, arith, stream)
        if gather:
            dist.all_gather_into_tensor(all_d, out_d)
            dist.all_gather_into_tensor(all_i, out_i)
            ndb.check(ndb._lib.load().ndb_b200_merge_topk_dev(ndb.ptr(all_d.data_ptr()), ndb.ptr(all_i.data_ptr()), world,
                                                              nq, k, ndb.ptr(fin_d.data_ptr()), ndb.ptr(fin_i.data_ptr()),
                                                              ndb.ptr(stream)))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(args.warmup):
        step(s)
    barrier()

    def burn(seconds):
        """Untimed steps of the same work: a step takes a fraction of a millisecond, so the K timed steps are over
        before nvidia-smi (100 ms period) samples once.  The timed region sits inside ~1 s of continuous identical
        load, and the clocks and throttle reasons reported are those of that second."""
        if gather:
            # the step contains collectives here: every rank must run the same number of them
            for s in range(int(seconds * 2048)):
                step(s)
            torch.cuda.synchronize()
            return
        t_end, s = time.perf_counter() + seconds, 0
        while time.perf_counter() < t_end:
            for _ in range(16):
                step(s)
                s += 1
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    burn(0.5)
    launches0 = ndb.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for s in range(args.steps):
        step(s)
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = ndb.launch_count() - launches0
    burn(0.5)
    if world > 1:
        t = torch.tensor([elapsed_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    clocks = sampler.stop()
    clocks["window"] = "0.5 s of untimed identical steps before and after the timed region"
    ms_per_step = elapsed_ms / args.steps
    value = nq * replicas / (ms_per_step * 1e-3)

    # recall@10 of the timed configuration against exact ground truth (first 500 queries of batch 0)
    step(0)
    torch.cuda.synchronize()
    res_i = (fin_i if gather else out_i)[:500].cpu().numpy()

    # dominant kernel (scan_topk_kernel): device time from CUDA events recorded by the library on the
    # launching stream, algorithmic bytes = sum over (query, probed list) of len*(dim*4+8)  
    ndb.set_timing(True)
    kms, kbytes, kevals = [], [], []
    for s in range(min(args.steps, 10)):
        step(s)
        ms, b, ev = ndb.last_kernel_stats()
        kms.append(ms); kbytes.append(b); kevals.append(ev)
    ndb.set_timing(False)
    kernel_ms = float(np.mean(kms))
    algo_bytes = float(np.mean(kbytes))
    evals = float(np.mean(kevals))

    # end to end through the host-pointer entry point with pinned host buffers
    qh = [torch.from_numpy(Q[i * nq:(i + 1) * nq]).pin_memory() for i in range(4)]
    hd = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    hi = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    lib = ndb._lib.load()

    def e2e_step(s, arith=arith):
        q = qh[s % 4]
        ndb.check(lib.ndb_b200_ivf_search(ix.h, ndb.ptr(q.data_ptr()), nq, w["nprobe"], k, ndb.IVF_FULL, arith,
                                          ndb.ptr(hd.data_ptr()), ndb.ptr(hi.data_ptr())))
        if gather:
            # ranks exchange their host results through the same NCCL path (device staging of 1.2 MB)
            out_d.copy_(hd, non_blocking=True); out_i.copy_(hi, non_blocking=True)
            dist.all_gather_into_tensor(all_d, out_d)
            dist.all_gather_into_tensor(all_i, out_i)
            ndb.check(lib.ndb_b200_merge_topk_dev(ndb.ptr(all_d.data_ptr()), ndb.ptr(all_i.data_ptr()), world, nq, k,
                                                  ndb.ptr(fin_d.data_ptr()), ndb.ptr(fin_i.data_ptr()), ndb.ptr(stream)))
            hd.copy_(fin_d, non_blocking=True); hi.copy_(fin_i, non_blocking=True)
            torch.cuda.synchronize()

    # pipelined form of the same call (ndb_b200_ivf_search_begin / _end, two batches in flight): the H2D
    # copy of the next batch and the D2H copy of the previous one overlap this batch's kernels.  Every
    # batch is still copied in from pinned host memory and its results copied back inside the timed region.
    hd2 = [hd.numpy(), torch.empty((nq, k), dtype=torch.float32).pin_memory().numpy()]
    hi2 = [hi.numpy(), torch.empty((nq, k), dtype=torch.int64).pin_memory().numpy()]
    qh_np = [q.numpy() for q in qh]

    def e2e_pipelined(nsteps, arith=arith):
        prev = None
        for s in range(nsteps):
            tk = ix.search_begin(qh_np[s % 4], hd2[s % 2], hi2[s % 2], w["nprobe"], k, ndb.IVF_FULL, arith)
            if prev is not None:
                ix.search_end(prev)
            prev = tk
        ix.search_end(prev)

    def timed(fn):
        barrier()
        t = time.perf_counter()
        fn()
        barrier()
        dt = (time.perf_counter() - t) / args.steps
        if world > 1:
            tt = torch.tensor([dt], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        return dt

    for s in range(3):
        e2e_step(s)
    e2e_sync_s = timed(lambda: [e2e_step(s) for s in range(args.steps)])
    if gather:
        e2e_s, e2e_mode = e2e_sync_s, "synchronous call per batch + NCCL all-gather and merge"
    else:
        e2e_pipelined(4)
        e2e_s = timed(lambda: e2e_pipelined(args.steps))
        e2e_mode = "search_begin/search_end, 2 batches in flight"
    e2e_val = nq * replicas / e2e_s

    # the same step in the reference's own fp32 arithmetic (bit-exact path), for comparison
    alt = None
    if args.arith == "tensor":
        a32 = ndb.ARITH_IVF_F32
        for s in range(3):
            step(s, a32)
        barrier()
        ev0.record()
        for s in range(args.steps):
            step(s, a32)
        ev1.record()
        barrier()
        alt_ms = ev0.elapsed_time(ev1) / args.steps
        for s in range(2):
            e2e_step(s, a32)
        barrier()
        t = time.perf_counter()
        for s in range(args.steps):
            e2e_step(s, a32)
        barrier()
        alt_e2e = (time.perf_counter() - t) / args.steps
        if world > 1:
            tt = torch.tensor([alt_ms, alt_e2e * 1e3], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            alt_ms, alt_e2e = float(tt[0].item()), float(tt[1].item()) * 1e-3
        step(0, a32)
        torch.cuda.synchronize()
        ref_i = (fin_i if gather else out_i)[:500].cpu().numpy()
        alt = {"arith": "ivf_f32", "value": nq * replicas / (alt_ms * 1e-3), "ms_per_step": alt_ms, "e2e": nq * replicas / alt_e2e,
               "unit": "queries/s", "ids_equal_to_tensor_path": float((ref_i == res_i).mean()),
               "note": "the reference's fp32 arithmetic end to end (bit-exact distances and ids vs the oracle)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    import workloads as W
    ngt = 500 if w["n"] <= 2_000_000 else 100
    gt = W.exact_ground_truth(X, Q[:ngt], k, w["metric"])
    recall = float(np.mean([len(set(a.tolist()) & set(b.tolist())) / k for a, b in zip(res_i[:ngt], gt)]))

    traffic = tensor_traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                tj = json.load(f).get(args.workload, {})
            traffic, tensor_traffic = tj.get("dram_bytes_per_launch"), tj.get("tensor_dram_bytes_per_launch")
        except Exception:
            traffic = tensor_traffic = None

    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    # the list-major kernel re-uses each list block for 8 queries from registers/L2, so its binding
    # limit is the FP32 pipe: 3 rounded ops (sub, mul, add) per element, 128 lanes per SM per clock
    fp32_ops = evals * dim * 3.0
    fp32_peak = 148 * 128 * sm_mhz * 1e6
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_kind,
                "kernel": "scan_topk_kernel (list mode)", "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_launch": algo_bytes, "distance_evals_per_launch": evals,
                "note": "algorithmic bytes = query-major definition of SURVEY 8d; the kernel groups queries per list "
                        "(8 per tile), so DRAM traffic is far below it and frac can exceed 1; see fp32",
                "fp32": {"achieved_tops": fp32_ops / (kernel_ms * 1e-3) / 1e12, "peak_tops": fp32_peak / 1e12,
                         "frac": fp32_ops / (kernel_ms * 1e-3) / fp32_peak, "unit": "T fp32 instr/s (non-fused)",
                         "sm_mhz": sm_mhz}}

    if args.arith == "tensor":
        # tc_knn_kernel (list mode): GEMM-form distances on tcgen05; algorithmic flops = 2 * dim per
        # (query, scanned vector) pair; the kernel runs inside a longer step -> sustained peak
        flops = 2.0 * evals * dim
        tpeak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        ach = flops / (kernel_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
                    "traffic": tensor_traffic, "peak_source": peak_kind, "kernel": "tc_knn_kernel (list mode)",
                    "kernel_ms": kernel_ms, "distance_evals_per_launch": evals,
                    "note": "algorithmic flops only: the kernel also multiplies tile padding (lists padded to 256 rows, "
                            "query groups padded to 128; about 7x the algorithmic flops on C2), which is not counted; the "
                            "kernel is bound by its top-k epilogue, see DESIGN.md 4.2b"}

    cpu = None
    if not args.no_cpu_baseline and world == 1:          # (rank 0 at N = 1 only: the host cores are shared by the ranks)
        import oracle_lib as O
        O.build_oracle()
        cores = os.cpu_count() or 1
        Cn = ix.centroids()
        sample = min(nq, 4000)
        qps_all, dt_all, _ = cpu_reference_qps(w, X, Q, Cn, sample, cores, native=True)
        qps_1, dt_1, _ = cpu_reference_qps(w, X, Q, Cn, min(500, sample), 1, native=True)
        # the reference's own build flags (-O2, no -march: NeuronDB/build.sh:712), all cores, smaller sample
        qps_o2, _, _ = cpu_reference_qps(w, X, Q, Cn, min(1000, sample), cores, native=False)
        cpu = {"value": qps_all, "unit": "queries/s", "cores": cores, "kind": "port",
               "sample": "%d queries of the 10k batch (%.1f s), oracle/ndb_oracle.c -O3 -march=native, OpenMP over "
                         "queries; 1 thread = %.0f QPS on %d queries; excludes PostgreSQL executor/bufmgr overhead"
                         % (sample, dt_all, qps_1, min(500, sample)),
               "value_1thread": qps_1, "value_reference_flags_O2": qps_o2}

    line = {
        "metric": "QPS@recall@10>=0.95", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if gather else "weak",
        "vs_baseline": None, "dtype": "bf16 select + f32 re-rank" if args.arith == "tensor" else "f32", "data": "synthetic",
        "config": {"workload": w["label"], "rows": w["n"], "dim": dim, "lists": w["lists"], "nprobe": w["nprobe"], "k": k,
                   "queries_per_step": nq * replicas, "arith": args.arith,
                   "l2": "inputs (%.2f GB of lists) larger than the 126 MB L2; 4 query batches rotate" % (w["n"] * dim * 4 / 1e9),
                   "parallelism": ("lists sharded l %% %d, queries replicated, NCCL all-gather + device merge" % world) if gather
                   else ("index replicated on %d GPUs, each answers its own %d-query batches (no data-path collective)"
                         % (world, nq)) if world > 1 else "1 GPU"},
        "recall_at_10": recall,
        "alt": alt,
        "e2e": {"value": e2e_val, "unit": "queries/s", "h2d_bytes_per_step": nq * replicas * dim * 4,
                "d2h_bytes_per_step": nq * replicas * k * 12, "ms_per_step": e2e_s * 1e3, "mode": e2e_mode,
                "synchronous_call": {"value": nq * replicas / e2e_sync_s, "ms_per_step": e2e_sync_s * 1e3}},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "build": build,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
