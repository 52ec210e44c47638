TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/multi_gpu_check.py > gpurun_out/mg2.json 2> gpurun_out/mg2.err; echo "mg2 rc=$?"
timeout 500 $TR bench.py --gpus 2 > gpurun_out/c4_n2.json 2> gpurun_out/c4_n2.err; echo "c4 n2 rc=$?"; tail -c 300 gpurun_out/c4_n2.err
