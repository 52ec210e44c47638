"""Print the few numbers of a bench.py JSON line that matter when comparing variants (reads stdin)."""
import json
import sys

for ln in sys.stdin:
    ln = ln.strip()
    if not ln.startswith("{"):
        continue
    d = json.loads(ln)
    r = d.get("roofline") or {}
    print("QPS %.3gM  step %.3f ms  recall %.3f  kernel %.3f ms  e2e %.3gM  launches %s  arith %s" % (
        d["value"] / 1e6, d["ms_per_step"], d.get("recall_at_10", -1), r.get("kernel_ms", -1), d["e2e"]["value"] / 1e6,
        d.get("gpu_launches"), d["config"].get("arith")))
