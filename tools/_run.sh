python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest.txt
python tools/c4_probe.py c4 > gpurun_out/probe_c4.txt 2>&1
python tools/c4_probe.py c2 > gpurun_out/probe_c2.txt 2>&1
python tools/tc_bench.py > gpurun_out/tc_bench_a.txt 2>&1
