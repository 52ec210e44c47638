python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest.txt
python tools/tc_bench.py > gpurun_out/tc_bench_a.txt 2>&1
NDB_TC_DENSE_NOSHARE=1 python tools/tc_bench.py > gpurun_out/tc_bench_b.txt 2>&1
NDB_TC_ITEMS_PER_SM=16 python tools/tc_bench.py > gpurun_out/tc_bench_c.txt 2>&1
NDB_TC_ITEMS_PER_SM=4 python tools/tc_bench.py > gpurun_out/tc_bench_d.txt 2>&1
python tools/tc_bench.py 2000000 10000 768 > gpurun_out/tc_bench_768.txt 2>&1
python bench.py --workload c3 --steps 5 --no-cpu-baseline --hnsw-efs 40 > gpurun_out/c3_hash.json 2> gpurun_out/c3_hash.err
NDB_HNSW_VISITED_GLOBAL=1 python bench.py --workload c3 --steps 5 --no-cpu-baseline --hnsw-efs 40 > gpurun_out/c3_global.json 2> gpurun_out/c3_global.err
