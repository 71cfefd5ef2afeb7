cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_aggregate.py -x -q -m gpu > gpurun_out/test_aggregate.log 2>&1; echo "rc=$?" >> gpurun_out/test_aggregate.log
tail -n 6 gpurun_out/test_aggregate.log | cut -c1-400
HYB_TRACE=1 timeout -k 10 600 python tools/variants.py --sf 10 --only aggregate > gpurun_out/variants.txt 2>&1
grep -v "^\[hyb\]" gpurun_out/variants.txt | tail -n 8; grep "shape" gpurun_out/variants.txt | sort | uniq -c
timeout -k 10 600 ncu --set full --import-source on --clock-control none -k regex:aggregate_stream -s 2 -c 1 -o gpurun_out/prof_r02e -f python tools/variants.py --sf 10 --only aggregate --repeat 3 > gpurun_out/ncu_agg.log 2>&1
tail -n 3 gpurun_out/ncu_agg.log | cut -c1-300
