cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_join.py -x -q -m gpu -k "peer_group or sorted_unique" > gpurun_out/test_peer.log 2>&1; echo "rc=$?" >> gpurun_out/test_peer.log
tail -n 25 gpurun_out/test_peer.log
timeout -k 10 600 python -m pytest tests/test_gpu_scan.py -x -q -m gpu -k "between" > gpurun_out/test_between.log 2>&1; echo "rc=$?" >> gpurun_out/test_between.log
tail -n 8 gpurun_out/test_between.log
timeout -k 10 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
tail -n 12 gpurun_out/smoke.log
