cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for f in q3 aggregate scan; do
  timeout -k 10 700 python -m pytest tests/test_gpu_$f.py -x -q -m gpu > gpurun_out/test_$f.log 2>&1; echo "rc=$?" >> gpurun_out/test_$f.log
  tail -n 30 gpurun_out/test_$f.log | cut -c1-600
done
