TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/multi_gpu_check.py > gpurun_out/mg2.json 2> gpurun_out/mg2.err; echo "mg2 rc=$?"; tail -1 gpurun_out/mg2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['all_ranks_ok'], d['exchange'])"; tail -3 gpurun_out/mg2.err
timeout 500 $TR bench.py --gpus 2 --steps 10 > gpurun_out/c4_n2.json 2> gpurun_out/c4_n2.err; echo "c4 n2 rc=$?"; tail -c 300 gpurun_out/c4_n2.err
