cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_join.py -x -q -m gpu > gpurun_out/test_join.log 2>&1; echo "rc=$?" >> gpurun_out/test_join.log
tail -n 25 gpurun_out/test_join.log
