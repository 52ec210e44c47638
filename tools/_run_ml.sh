# round-2 late: parity of the quantised paths, bench lines, ncu captures of their kernels
python -m pytest tests/test_gpu_pq.py -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_pq.txt
python bench.py --workload pq --steps 10 > gpurun_out/bench_pq.json 2> gpurun_out/bench_pq.err
NDB_PQ_NO_FILTER=1 python bench.py --workload pq --steps 10 --no-cpu-baseline > gpurun_out/bench_pq_nofilter.json 2> gpurun_out/bench_pq_nofilter.err
for part in pq ham ckm; do
  case $part in pq) kern=pq_adc_kernel;; ham) kern=hamming_topk_kernel;; ckm) kern=ckm_assign_kernel;; esac
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:$kern --launch-count 1 -f -o gpurun_out/r02e_$part python tools/ml_probe.py $part > gpurun_out/ncu_$part.log 2>&1
done
timeout 200 python tools/ml_bench.py > gpurun_out/ml_bench.jsonl 2> gpurun_out/ml_bench.err
cat gpurun_out/pytest_pq.txt; cut -c1-1500 gpurun_out/bench_pq.json; tail -3 gpurun_out/bench_pq.err; cut -c1-400 gpurun_out/bench_pq_nofilter.json; tail -2 gpurun_out/ncu_*.log; grep -h pq_scan gpurun_out/ml_bench.jsonl | cut -c1-500
