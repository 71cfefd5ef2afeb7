cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --import-source on --clock-control none -k regex:aggregate_stream -s 2 -c 1 -o gpurun_out/prof_r02d -f python tools/variants.py --sf 10 --only aggregate --repeat 3 > gpurun_out/ncu_agg.log 2>&1
tail -n 5 gpurun_out/ncu_agg.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
