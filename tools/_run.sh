python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/pytest.txt
python tools/c4_probe.py c4 > gpurun_out/probe_c4.txt 2>&1
