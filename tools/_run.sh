python -m pytest tests/test_gpu_tensor.py -m gpu -x -q 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:"tc_knn_kernel" --launch-skip 3 --launch-count 1 -o gpurun_out/c4_list_full -f python tools/cert_stats.py c4 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/cert_c2_launches.csv python tools/cert_stats.py c2 > /dev/null 2>&1
