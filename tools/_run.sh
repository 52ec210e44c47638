python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest.txt
python bench.py --steps 10 > gpurun_out/b_full.json 2> gpurun_out/b_full.err; tail -c 300 gpurun_out/b_full.err
