cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_join.py -x -q -m gpu -k "peer_group" > gpurun_out/test_peer.log 2>&1; echo "rc=$?" >> gpurun_out/test_peer.log
tail -n 12 gpurun_out/test_peer.log | cut -c1-800
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tests/gpu_distributed_worker.py > gpurun_out/worker2.log 2>&1; echo "rc=$?" >> gpurun_out/worker2.log
tail -n 12 gpurun_out/worker2.log | cut -c1-900
