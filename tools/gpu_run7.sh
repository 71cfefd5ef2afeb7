cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 600 python tools/variants.py --sf 10 > gpurun_out/variants.txt 2>&1
grep "scan\|aggregate" gpurun_out/variants.txt
timeout -k 10 1500 ncu --set full --clock-control none --import-source on -k regex:'aggregate_stream|join_span_count|join_build' -s 4 -c 3 -o gpurun_out/prof_r02c python bench.py --sf 10 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
