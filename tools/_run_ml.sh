# round-2 late: parity of the quantised / ML paths, bench lines, ncu captures of their kernels
python -m pytest tests/test_gpu_quant.py tests/test_gpu_pq.py tests/test_gpu_mlkmeans.py tests/test_gpu_mlknn.py tests/test_gpu_boundary.py -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_ml.txt
python bench.py --workload pq --steps 10 > gpurun_out/bench_pq.json 2> gpurun_out/bench_pq.err
for part in pq ham ckm; do
  case $part in pq) kern=pq_adc_kernel;; ham) kern=hamming_topk_mq_kernel;; ckm) kern=ckm_assign_kernel;; esac
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:$kern --launch-count 1 -f -o gpurun_out/r02e_$part python tools/ml_probe.py $part > gpurun_out/ncu_$part.log 2>&1
done
timeout 200 python tools/ml_bench.py > gpurun_out/ml_bench.jsonl 2> gpurun_out/ml_bench.err
cat gpurun_out/pytest_ml.txt; cut -c1-300 gpurun_out/bench_pq.json; tail -n 3 gpurun_out/bench_pq.err; for f in gpurun_out/ncu_*.log; do tail -n 2 $f; done; cut -c1-330 gpurun_out/ml_bench.jsonl; tail -n 3 gpurun_out/ml_bench.err
