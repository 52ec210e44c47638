python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest.txt
python tools/c4_probe.py c4 > gpurun_out/probe_c4.txt 2>&1
PROBE_STRIPE=8 python tools/c4_probe.py c4 > gpurun_out/probe_c4_s8.txt 2>&1
python tools/c4_probe.py c2 > gpurun_out/probe_c2.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.txt 2>&1
