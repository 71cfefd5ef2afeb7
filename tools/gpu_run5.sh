cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 1500 ncu --set full --clock-control none --import-source on -k regex:'scan_bulk|aggregate_stream|join_span|join_build' -s 8 -c 5 -o gpurun_out/prof_r02b python bench.py --sf 10 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_full.log | head -c 600
