"""Development aid: build one IVF workload (default C4) once, then time the tensor-path search under several
environment-variable variants (the library reads them per call).  Prints one line per variant: step ms,
list-kernel ms, result equality against the first variant and, when the library was built with
-DNDB_TC_COUNTERS, the epilogue statistics of the list kernel.

    python tools/c4_probe.py [c4|c2] "VAR=1 VAR2=x" "VAR3=y" ...
"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bench
import neurondb_b200 as ndb
from neurondb_b200 import _lib


def main():
    args = sys.argv[1:]
    wname = "c4"
    if args and args[0] in bench.WORKLOADS:
        wname = args.pop(0)
    variants = [""] + args
    w = bench.WORKLOADS[wname]
    torch.cuda.set_device(0)
    ndb.init(0)
    X, Q = bench.make_data(w)
    n, nq, k = w["n"], w["nq"], w["k"]
    ix = ndb.IvfIndex(w["dim"], w["lists"], w["metric"])
    t0 = time.perf_counter()
    ix.ivfbuild(X)
    stripe = int(os.environ.get("PROBE_STRIPE", "1"))      # one rank's share of a striped N-GPU index (row i on rank i % N)
    if stripe > 1:
        ix.ivfinsert(X[0::stripe], np.arange(0, n, stripe, dtype=np.int64))
    else:
        ix.ivfinsert(X, np.arange(n, dtype=np.int64))
    ix.prepare(ndb.ARITH_TENSOR)
    print(f"{wname}: build {time.perf_counter() - t0:.2f} s", flush=True)
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    qd = [torch.from_numpy(Q[i * nq:(i + 1) * nq]).cuda() for i in range(4)]
    out_d = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    out_i = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    lib = _lib.load()
    have_ctr = hasattr(lib, "ndbdbg_tc_counters")
    try:
        lib.ndbdbg_tc_counters.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
    except AttributeError:
        have_ctr = False

    def counters():
        buf = (ctypes.c_ulonglong * 16)()
        lib.ndbdbg_tc_counters(buf)
        return list(buf)

    def step(b):
        ix.search_dev(qd[b].data_ptr(), nq, out_d.data_ptr(), out_i.data_ptr(), w["nprobe"], k, ndb.IVF_FULL, ndb.ARITH_TENSOR,
                      st.cuda_stream)

    ref = None
    for v in variants:
        sets = dict(kv.split("=", 1) for kv in v.split()) if v else {}
        one_batch = sets.pop("ONE_BATCH", None)
        for kk, vv in sets.items():
            os.environ[kk] = vv
        for s in range(3):
            step(0 if one_batch else s % 4)
        torch.cuda.synchronize()
        if os.environ.get("PROBE_PROFILE"):          # under `ncu --profile-from-start off`: capture exactly one step
            torch.cuda.profiler.start()
            step(0)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        if have_ctr:
            counters()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 8
        ev0.record()
        for s in range(steps):
            step(0 if one_batch else s % 4)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / steps
        ctr = counters() if have_ctr else None
        ndb.set_timing(True)
        km = []
        for s in range(4):
            step(0 if one_batch else s % 4)
            km.append(ndb.last_kernel_stats()[0])
        ndb.set_timing(False)
        step(0)
        torch.cuda.synchronize()
        ri, rd = out_i.cpu().numpy().copy(), out_d.cpu().numpy().copy()
        if ref is None:
            ref = (ri, rd)
        same = bool((ri == ref[0]).all() and (rd.view(np.uint32) == ref[1].view(np.uint32)).all())
        cs = ix.cert_stats()
        line = f"[{v or 'default'}] step {ms:.3f} ms  list kernel {np.mean(km):.3f} ms  same_as_default {same}  fallback_q {cs['list_fallback_queries']} exact_evals/q {cs['list_exact_evals'] / nq:.1f}"
        if ctr:
            for name, off in (("phase 1 / single", 0), ("phase 2", 8)):
                ch, anyc, heavy, iters, takers = [c / steps for c in ctr[off:off + 5]]
                if ch:
                    line += (f"\n    {name}: warp-chunks {ch:.3g} any {anyc / max(ch, 1):.3f} heavy {heavy / max(ch, 1):.4f} insert-rounds/chunk "
                             f"{iters / max(ch, 1):.3f} takers/query {takers / nq:.0f}")
        print(line, flush=True)
        for kk in sets:
            os.environ.pop(kk, None)


if __name__ == "__main__":
    main()
