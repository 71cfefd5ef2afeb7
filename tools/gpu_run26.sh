cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_scan.py tests/test_gpu_aggregate.py -x -q -m gpu > gpurun_out/test_scan.log 2>&1; echo "rc=$?" >> gpurun_out/test_scan.log
tail -n 5 gpurun_out/test_scan.log | cut -c1-600
timeout -k 10 600 python tools/variants.py --sf 10 --only scan > gpurun_out/variants.txt 2>&1
tail -n 4 gpurun_out/variants.txt
timeout -k 10 600 python tools/variants.py --sf 100 --only scan --repeat 3 > gpurun_out/variants100.txt 2>&1
tail -n 4 gpurun_out/variants100.txt
timeout -k 10 600 python tools/step_gaps.py --sf 100 > gpurun_out/gaps.txt 2>&1
cat gpurun_out/gaps.txt | tail -n 10
