cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/host.txt; nproc >> gpurun_out/host.txt
( time timeout -k 10 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_sf100.json 2> gpurun_out/bench_sf100.err ) 2> gpurun_out/bench_sf100.time; echo "rc=$?" >> gpurun_out/bench_sf100.err
cat gpurun_out/host.txt gpurun_out/bench_sf100.time; tail -5 gpurun_out/bench_sf100.err; head -c 6000 gpurun_out/bench_sf100.json
