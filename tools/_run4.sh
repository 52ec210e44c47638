TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
timeout 400 $TR4 bench.py --gpus 4 > gpurun_out/c4_n4.json 2> gpurun_out/c4_n4.err; echo "c4 n4 rc=$?"; tail -c 300 gpurun_out/c4_n4.err
