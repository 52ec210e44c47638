#!/usr/bin/env python
"""Summarise ncu output into small text files for profiles/ (the .ncu-rep files are scratch).

  ncu_summary.py launches <launches.csv>           -> per-kernel launch count / total time / share
  ncu_summary.py steps    <launches.csv> <kernel>  -> per-step sequence, a step ending with <kernel>
  ncu_summary.py kernel   <file.ncu-rep> [index]   -> key metrics, stall reasons, SASS opcode mix
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def steps(path, last_kernel):
    """Per-step view of a launch list: a step is the run of launches that ends with `last_kernel`
    (e.g. ivf_tc_finish_kernel); prints the mean duration of each position over the steps found."""
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [(re.sub(r"\(.*", "", r["Kernel Name"])[:80], float(r["Metric Value"].replace(",", "")) / 1000.0)
            for r in csv.DictReader(lines)]
    ends = [i for i, (n, _) in enumerate(rows) if last_kernel in n]
    seqs = []
    for a, b in zip(ends, ends[1:]):
        seq = rows[a + 1:b + 1]
        if len(seq) <= 16:
            seqs.append(seq)
    if not seqs:
        print("no steps found")
        return
    shape = collections.Counter(tuple(n for n, _ in q) for q in seqs).most_common(1)[0][0]
    seqs = [q for q in seqs if tuple(n for n, _ in q) == shape]
    print("# one step = %d launches, mean over %d steps (ncu gpu__time_duration.sum: cold-cache, serialised)" % (len(shape), len(seqs)))
    tot = 0.0
    for i, name in enumerate(shape):
        us = sum(q[i][1] for q in seqs) / len(seqs)
        tot += us
        print("%9.1f us  %s" % (us, name))
    print("%9.1f us  total" % tot)


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:90]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# per-kernel device time (ncu --metrics gpu__time_duration.sum, cold-cache, serialised)")
    print("%12s %6s %7s  %s" % ("total_us", "count", "share", "kernel"))
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%12.1f %6d %6.1f%%  %s" % (t, c, 100 * t / tot, n))


def kernel(path, index=0):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2 + index]
    print("# kernel:", r[hdr.index("Kernel Name")][:120])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("%-72s %18s %s" % (k, r[i], units[i]))
    print("# warp stall reasons (cycles per issued instruction)")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            print("  %-28s %s" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), r[i]))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks = []
    for b in src.split('"Kernel Name",')[1:]:       # (--import-source on: most kernels appear twice in a row, as two views)
        if not blocks or b != blocks[-1]:
            blocks.append(b)
    if index < len(blocks):
        lines = blocks[index].split("\n")
        srows = list(csv.reader(lines[1:]))
        h = srows[0]
        ia, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        data = [x for x in srows[1:] if len(x) == len(h)]
        tot = sum(int(x[isamp] or 0) for x in data)
        totex = sum(int(x[iex] or 0) for x in data)
        ops, samp = collections.Counter(), collections.Counter()
        for x in data:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", x[ia])
            op = m.group(2).split(".")[0] if m else "?"
            ops[op] += int(x[iex] or 0)
            samp[op] += int(x[isamp] or 0)
        print("# SASS opcode mix (share of executed warp instructions / of stall samples)")
        for op, c in ops.most_common(16):
            print("  %-10s exec %5.1f%%  samples %5.1f%%" % (op, 100 * c / max(totex, 1), 100 * samp[op] / max(tot, 1)))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "steps":
        steps(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
