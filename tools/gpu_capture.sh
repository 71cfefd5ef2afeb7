# The ncu evidence of a round (single B200, about 6 GPU-minutes): launch list of the bench command, one full-set capture of a
# step's kernels, then feed both to tools/make_profiles.py on the CPU box.
#   /usr/local/graft/bin/gpurun --timeout 3600 -- 'bash tools/gpu_capture.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout -k 10 1800 ncu --set full --clock-control none --import-source on -k regex:'scan_mask|scan_bases|scan_expand|aggregate_stream|join_span|join_build|exclusive_scan' -s 24 -c 8 -o gpurun_out/prof_step -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/ncu_full.log 2>&1; echo "full capture rc=$?"
ls -la gpurun_out/prof_step.ncu-rep gpurun_out/launches.csv
