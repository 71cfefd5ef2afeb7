# bench.py on N GPUs of one box, as the driver launches it:
#   /usr/local/graft/bin/gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_scaling.sh N'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29640 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
tail -c 3000 gpurun_out/bench_n$N.json
