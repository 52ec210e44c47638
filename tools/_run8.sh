TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/multi_gpu_check.py > gpurun_out/mg8.json 2> gpurun_out/mg8.err; echo "mg8 rc=$?"; tail -1 gpurun_out/mg8.json | cut -c1-300
timeout 500 $TR bench.py --gpus 8 > gpurun_out/c4_n8.json 2> gpurun_out/c4_n8.err; echo "c4 n8 rc=$?"; tail -c 300 gpurun_out/c4_n8.err
timeout 500 $TR bench.py --gpus 8 --workload c5 --steps 5 > gpurun_out/c5_n8.json 2> gpurun_out/c5_n8.err; echo "c5 n8 rc=$?"; tail -c 300 gpurun_out/c5_n8.err
