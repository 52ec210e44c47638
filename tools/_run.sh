python bench.py --workload c5 --steps 5 --no-cpu-baseline > gpurun_out/c5_n1.json 2> gpurun_out/c5_n1.err
python tools/c4_probe.py c2 > gpurun_out/probe_c2.txt 2>&1
python tools/c4_probe.py c4 > gpurun_out/probe_c4.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest.txt
