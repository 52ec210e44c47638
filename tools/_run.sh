python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --no-cpu-baseline > gpurun_out/b_np4.json 2> gpurun_out/b_np4.err; tail -c 300 gpurun_out/b_np4.err
