python bench.py > gpurun_out/r02_bench_c4_n1_with_c2.json 2> gpurun_out/b_full.err; tail -c 300 gpurun_out/b_full.err
python bench.py --workload c1 > gpurun_out/r02_bench_c1_n1.json 2> gpurun_out/c1.err
python bench.py --workload c3 > gpurun_out/r02_bench_c3_n1.json 2> gpurun_out/c3.err
python bench.py --workload c5 > gpurun_out/r02_bench_c5_n1.json 2> gpurun_out/c5.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/ref.err
