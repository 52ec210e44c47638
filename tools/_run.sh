python -m pytest tests/test_gpu_ivf.py tests/test_gpu_core.py tests/test_gpu_tensor.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 10 --no-cpu-baseline > gpurun_out/b_np4.json 2> gpurun_out/b_np4.err; tail -c 300 gpurun_out/b_np4.err
python bench.py --workload c5s --steps 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['kmeans'], d['value'])"
