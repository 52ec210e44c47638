python -m pytest tests/test_gpu_tensor.py -m gpu -x -q 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_block_queries|ivf_tc_finish|ivf_exact_fallback" --csv --log-file gpurun_out/cert_c4_launches.csv python tools/cert_stats.py c4 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_block_queries|ivf_tc_finish|ivf_exact_fallback" --csv --log-file gpurun_out/cert_c2_launches.csv python tools/cert_stats.py c2 > /dev/null 2>&1
