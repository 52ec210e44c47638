#!/usr/bin/env python
"""Compile one .cu with -Xptxas -v and print a compact per-kernel register/spill/smem table."""
import re, subprocess, sys
src = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
       "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++", "--expt-relaxed-constexpr", "-Xptxas", "-v", "-c", src,
       "-o", "/tmp/ptxas_report.o"]
out = subprocess.run(cmd, capture_output=True, text=True).stderr
name = None
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*$", "", name)
        spill = ""
        continue
    m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and name:
        spill = f"spill {m.group(1)}/{m.group(2)}"
    m = re.search(r"Used (\d+) registers.*?(?:, (\d+) bytes smem)?", line)
    if m and name:
        sm = re.search(r"(\d+) bytes smem", line)
        if pat in name:
            print(f"{m.group(1):>4} regs  {spill:14s} smem {sm.group(1) if sm else 0:>6}  {name}")
        name = None
if "error" in out:
    print(out[-3000:])
