cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 700 python -m pytest tests/test_gpu_aggregate.py -x -q -m gpu > gpurun_out/test_aggregate.log 2>&1; echo "rc=$?" >> gpurun_out/test_aggregate.log
tail -n 3 gpurun_out/test_aggregate.log
timeout -k 10 600 python tools/variants.py --sf 10 --only scan,aggregate > gpurun_out/variants.txt 2>&1
cat gpurun_out/variants.txt
timeout -k 10 600 python tools/variants.py --sf 100 --only scan > gpurun_out/variants100.txt 2>&1
cat gpurun_out/variants100.txt
