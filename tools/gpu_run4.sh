cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for f in scan join aggregate; do
  timeout -k 10 700 python -m pytest tests/test_gpu_$f.py -x -q -m gpu > gpurun_out/test_$f.log 2>&1; echo "rc=$?" >> gpurun_out/test_$f.log
  tail -n 4 gpurun_out/test_$f.log
done
timeout -k 10 600 python tools/variants.py --sf 10 > gpurun_out/variants.txt 2>&1
cat gpurun_out/variants.txt
