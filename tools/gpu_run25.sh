cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_join.py tests/test_cpp_mirror.py tests/test_gpu_q3.py tests/test_gpu_distributed.py -x -q -m gpu > gpurun_out/test_join.log 2>&1; echo "rc=$?" >> gpurun_out/test_join.log
tail -n 5 gpurun_out/test_join.log | cut -c1-600
timeout -k 10 600 python tools/step_gaps.py --sf 100 > gpurun_out/gaps.txt 2>&1
cat gpurun_out/gaps.txt | tail -n 14
