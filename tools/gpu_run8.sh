cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 700 python -m pytest tests/test_gpu_aggregate.py -x -q -m gpu > gpurun_out/test_aggregate.log 2>&1; echo "rc=$?" >> gpurun_out/test_aggregate.log
tail -n 4 gpurun_out/test_aggregate.log
timeout -k 10 600 python tools/variants.py --sf 10 > gpurun_out/variants.txt 2>&1
grep "aggregate\|auto" gpurun_out/variants.txt
