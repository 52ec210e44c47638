TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/multi_gpu_check.py > gpurun_out/mg2.json 2> gpurun_out/mg2.err; echo "mg2 rc=$?"; tail -1 gpurun_out/mg2.json | cut -c1-300
timeout 500 $TR bench.py --gpus 2 --steps 10 > gpurun_out/c4_n2.json 2> gpurun_out/c4_n2.err; echo "c4 n2 rc=$?"; tail -c 300 gpurun_out/c4_n2.err
timeout 600 python bench.py --workload c3 --steps 5 > gpurun_out/c3_full.json 2> gpurun_out/c3_full.err; echo "c3 rc=$?"; tail -c 300 gpurun_out/c3_full.err
timeout 400 python bench.py --workload c5 --steps 5 > gpurun_out/c5_n1.json 2> gpurun_out/c5_n1.err; echo "c5 rc=$?"; tail -c 300 gpurun_out/c5_n1.err
timeout 300 python bench.py --workload c1 --steps 10 > gpurun_out/c1_n1.json 2> gpurun_out/c1_n1.err; echo "c1 rc=$?"
