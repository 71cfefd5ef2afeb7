cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_scan.py tests/test_gpu_join.py tests/test_gpu_q3.py -x -q -m gpu > gpurun_out/test_scan.log 2>&1; echo "rc=$?" >> gpurun_out/test_scan.log
tail -n 4 gpurun_out/test_scan.log | cut -c1-600
timeout -k 10 600 python tools/variants.py --sf 10 --only scan,join > gpurun_out/variants.txt 2>&1
tail -n 9 gpurun_out/variants.txt
timeout -k 10 600 python tools/variants.py --sf 100 --only scan,join --repeat 3 > gpurun_out/variants100.txt 2>&1
tail -n 9 gpurun_out/variants100.txt
