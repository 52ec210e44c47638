TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for w in c1 c5s; do timeout 300 python bench.py --workload $w --steps 5 > gpurun_out/w_$w.json 2> gpurun_out/w_$w.err; echo "$w rc=$?"; tail -c 300 gpurun_out/w_$w.err; done
NDB_BENCH_C3_ROWS=100000 timeout 400 python bench.py --workload c3 --steps 5 > gpurun_out/w_c3.json 2> gpurun_out/w_c3.err; echo "c3 rc=$?"; tail -c 300 gpurun_out/w_c3.err
timeout 300 $TR bench.py --gpus 2 --workload c5s --steps 5 > gpurun_out/w_c5s_n2.json 2> gpurun_out/w_c5s_n2.err; echo "c5s n2 rc=$?"; tail -c 300 gpurun_out/w_c5s_n2.err
NDB_BENCH_C3_ROWS=100000 timeout 400 $TR bench.py --gpus 2 --workload c3 --steps 5 > gpurun_out/w_c3_n2.json 2> gpurun_out/w_c3_n2.err; echo "c3 n2 rc=$?"; tail -c 300 gpurun_out/w_c3_n2.err
timeout 300 $TR bench.py --gpus 2 --workload c1 --steps 5 > gpurun_out/w_c1_n2.json 2> gpurun_out/w_c1_n2.err; echo "c1 n2 rc=$?"; tail -c 300 gpurun_out/w_c1_n2.err
