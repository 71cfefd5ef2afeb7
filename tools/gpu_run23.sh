cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tests/gpu_distributed_worker.py > gpurun_out/worker2.log 2>&1; echo "rc=$?" >> gpurun_out/worker2.log
tail -n 12 gpurun_out/worker2.log | cut -c1-900
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"
tail -c 6000 gpurun_out/bench_n2.json
tail -n 5 gpurun_out/bench_n2.err | cut -c1-400
