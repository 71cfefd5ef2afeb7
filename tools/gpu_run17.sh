cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 700 python -m pytest tests/test_gpu_q3.py -x -q -m gpu > gpurun_out/test_q3.log 2>&1; echo "rc=$?" >> gpurun_out/test_q3.log
tail -n 30 gpurun_out/test_q3.log | cut -c1-700
