cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 600 python tools/variants.py --sf 10 --only aggregate > gpurun_out/variants.txt 2>&1
tail -n 8 gpurun_out/variants.txt
timeout -k 10 300 python -m pytest tests/test_gpu_aggregate.py -x -q -m gpu 2>&1 | tail -n 3
