cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
./tools/micro/gather > gpurun_out/gather.txt 2>&1
for f in scan join aggregate; do
  timeout -k 10 700 python -m pytest tests/test_gpu_$f.py -x -q -m gpu > gpurun_out/test_$f.log 2>&1; echo "rc=$?" >> gpurun_out/test_$f.log
done
timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "rc=$?" >> gpurun_out/bench1.err
tail -3 gpurun_out/test_*.log; cat gpurun_out/gather.txt; cat gpurun_out/bench1.json | head -c 3000
