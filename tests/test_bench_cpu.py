"""CPU: the host-side pieces of bench.py that do not need a GPU (workload table, peaks fallback,
nvidia-smi line parsing, the reference arm's JSON contract on the smoke workload)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_workloads_follow_baseline_configs():
    c2 = bench.WORKLOADS["c2"]
    assert (c2["n"], c2["dim"], c2["lists"], c2["nprobe"], c2["k"], c2["nq"], c2["metric"]) == (1_000_000, 128, 1024, 16, 10, 10_000, 1)
    c4 = bench.WORKLOADS["c4"]
    assert (c4["n"], c4["dim"], c4["lists"], c4["metric"]) == (10_000_000, 96, 4096, 3)
    X, Q = bench.make_data(bench.WORKLOADS["smoke"])
    assert X.shape == (50_000, 64) and Q.shape == (4_000, 64) and X.dtype == np.float32
    X2, Q2 = bench.make_data(bench.WORKLOADS["smoke"], qstream=1)        # another replica's queries, same rows
    assert np.array_equal(X, X2) and not np.array_equal(Q, Q2)


def test_clock_sampler_parses_nvidia_smi_lines():
    s = bench.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None})()
    s.lines = ["0, 1410, 1965, 310.5, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active",
               "0, 1965, 1965, 720.1, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
               "0, 1965, 1965, 731.0, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active",
               "garbage"]
    c = s.stop()
    assert c["sm_mhz"] == 1965.0 and c["sm_max_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"] and c["samples"] == 3


def test_measured_peaks_has_the_keys_the_roofline_needs():
    peaks, kind = bench.measured_peaks()
    assert kind in ("measured", "fallback") and peaks["hbm_gbs"] > 1000 and peaks["bf16_tflops"] > 100


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "smoke",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "QPS@recall@10>=0.95" and line["unit"] == "queries/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
